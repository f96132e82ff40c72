# A/B of the compute-bound (strided / state-only) kernels; prints the mode-kernel lines of dev_probe
set -u
run() { echo "== $*"; env "$@" | grep -E "^K2" | tail -2; }
for mb in 4 5 6 8; do run PF_FO1_MINB=$mb PF_EVERY=8163 python scripts/dev_probe.py c5 1048576; done
run PF_FO2_COMPLEX=1 PF_EVERY=64 python scripts/dev_probe.py c3 1048576
for mb in 4 6 8; do run PF_FO2_MINB=$mb PF_EVERY=64 python scripts/dev_probe.py c3 1048576; done
run PF_EVERY=512 python scripts/dev_probe.py c4 65536
run python scripts/dev_probe.py c4 8192
