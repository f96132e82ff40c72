# mode-kernel lines of dev_probe for the compute-bound (strided / state-only) launches and the
# full-trajectory ones (records already built)
set -u
run() { echo "== $*"; env "$@" | grep -E "^K2|^WINDOW" | tail -2; }
run PF_EVERY=8163 python scripts/dev_probe.py c5 1048576
run PF_EVERY=64 python scripts/dev_probe.py c3 1048576
run PF_EVERY=512 python scripts/dev_probe.py c4 65536
run python scripts/dev_probe.py c4 8192
run python scripts/dev_probe.py c3 65536
run python scripts/dev_probe.py c2
