# Exposure map of the captured step: ms/step with whole C-ABI calls dropped (debug library only; results are INVALID, timing only).
export FFPN_LIB=debug
run() { FFPN_TIMING_SKIP=$1 python tools/ab_step.py 30 > gpurun_out/ab_tmp.log 2>&1; grep "^AB " gpurun_out/ab_tmp.log | sed 's/ B8 S32.*slo 320x128//; s/ loss.*//' || tail -5 gpurun_out/ab_tmp.log; }
for s in "" ffpn_conv_wgrad ffpn_conv_dgrad ffpn_block_end_bwd ffpn_block_end_fwd ffpn_bn_bwd_apply ffpn_bn_bwd_reduce ffpn_conv_wgrad,ffpn_conv_dgrad,ffpn_block_end_bwd,ffpn_bn_bwd_apply,ffpn_bn_bwd_reduce,ffpn_bn_bwd_finalize,ffpn_maxpool_bwd,ffpn_proj_tail_bwd,ffpn_resize2d_bwd,ffpn_upsample_bwd,ffpn_head_bwd; do run "$s"; done
