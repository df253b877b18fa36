# usage: bash scripts/exp_run.sh <logname> ; reads experiment lines "CFG | ENV" from scripts/exp_list.txt
LOG=gpurun_out/$1.log; : > $LOG
while IFS='|' read -r CFG ENVS; do
  [ -z "$CFG$ENVS" ] && continue
  echo "== cfg:[$CFG] env:[$ENVS]" >> $LOG
  env $ENVS python bench.py $CFG --nsteps 120 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python scripts/exp_fmt.py >> $LOG
done < scripts/exp_list.txt
cat $LOG
