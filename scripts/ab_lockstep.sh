# A/B of the persistent lock-step launches (full-trajectory, records already built)
set -u
run() { echo "== $*"; env "$@" | grep -E "^K2|^WINDOW" | tail -2; }
run python scripts/dev_probe.py c2
run PF_FO1_NOLOCKSTEP=1 python scripts/dev_probe.py c2
run PF_WINDOW=3000,4024 python scripts/dev_probe.py c5 1048576
run PF_FO1_NOLOCKSTEP=1 PF_WINDOW=3000,4024 python scripts/dev_probe.py c5 1048576
run() { echo "== $*"; env "$@" | grep -E "^SOLVE" | tail -2; }
run python scripts/dev_probe.py c2 131072
run PF_FO1_NOLOCKSTEP=1 python scripts/dev_probe.py c2 131072
