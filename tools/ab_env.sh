# ms/step of the captured step under a list of environment settings: bash tools/ab_env.sh "A=1 B=2" "A=3" ...
for e in "$@"; do
  env $e timeout 100 python tools/ab_step.py 30 2>&1 | grep "^AB " | sed 's/ B8 S32.*slo 320x128//; s/ loss.*//'
done
