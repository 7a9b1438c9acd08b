# GPU-box helper: group-size experiment (SGC_LS_GK variants in build/variants): bench each, parity tests on gk32 and on the default build
bash scripts/run_variants.sh
cp spacegraphcats_b200/libsgc_b200.so /tmp/base.so
cp build/variants/gk32.so spacegraphcats_b200/libsgc_b200.so
echo "tests on gk32:"; timeout 300 python -m pytest tests -m gpu -x -q -k "label or hypothesis or full_size or index_reads or count_matches or batched or peer_probed or smoke" 2>&1 | tail -2
cp /tmp/base.so spacegraphcats_b200/libsgc_b200.so
echo "tests on the default build:"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
