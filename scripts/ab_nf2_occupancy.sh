run() { echo "== $*"; env "$@" | grep -E "^K2|^SOLVE" | tail -3; }
run PF_EVERY=64 python scripts/dev_probe.py c3 1048576
for mb in 4 5 6; do run PYFLATION_B200_LIB=/root/repo/pyflation_b200/lib/alt_nf2_mb$mb.so PF_EVERY=64 python scripts/dev_probe.py c3 1048576; done
run python scripts/dev_probe.py c3 65536
for mb in 4 5 6; do run PYFLATION_B200_LIB=/root/repo/pyflation_b200/lib/alt_nf2_mb$mb.so python scripts/dev_probe.py c3 65536; done
