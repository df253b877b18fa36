# sourced by the ci_*.sh ports: the reference's scripts/ci_*.sh flows (klindsay28/Newton-Krylov_OOC) on the B200 path.
#   python -m nk_ooc.<model>.setup_solver / nk_driver.sh  ->  python -m nk_ooc_b200.cli setup_solver / nk_driver
#   python -m nk_ooc.baseline_cmp                         ->  python -m nk_ooc_b200.baseline_cmp
# Same work directories ($HOME/ci_*_workdir), file lists and tolerances as the reference's scripts.  The
# reference persists its configuration in the work directory (--persist) and nk_driver.sh reads it back; here the
# same options are passed to both commands.  NKB_BASELINES = a copy of the reference's baselines/ directory
# (tests/baseline_files.py builds one from tests/golden on the GPU box).
root=$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)
export PYTHONPATH=$root/newton-krylov_ooc_b200${PYTHONPATH:+:$PYTHONPATH}
baselines=${NKB_BASELINES:-$root/baselines}
cli="python -m nk_ooc_b200.cli"
err_cnt=0

baseline_cmp() {  # baseline_cmp <expr_dir> <baseline_dir> <fname> [--rtol R] [--atol A]
    local expr_dir=$1 baseline_dir=$2 fname=$3
    shift 3
    echo comparing $fname
    python -m nk_ooc_b200.baseline_cmp --fname $fname --expr_dir $expr_dir --baseline_dir $baseline_dir "$@" \
        || err_cnt=$((err_cnt+1))
}

newton_state_cmp() {  # newton_state_cmp <expr_dir> <baseline_dir>
    echo comparing Newton_state.json to baseline
    diff -u -b <(sed "s%$HOME%HOME%g" $1/Newton_state.json) $2/Newton_state.json || err_cnt=$((err_cnt+1))
}
