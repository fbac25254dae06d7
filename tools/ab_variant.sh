# Kernel-case timings (eager, CUDA events) + the captured step; FFPN_LIB=debug selects the build parked in the debug slot.
for c in ${CASES:-conv_fwd_l1 conv_dgrad_l1 conv_dgrad_bnr_l1 conv_dgrad_bnr_l2 conv_dgrad_bnr_l3}; do
  timeout 60 python tools/kernel_cases.py $c 30 | tail -1 | cut -c1-60
done
timeout 100 python tools/ab_step.py 30 2>&1 | grep "^AB " | sed 's/ B8 S32.*slo 320x128//; s/ loss.*//'
