#!/bin/bash
# warp-specialised C1 kernel: quick correctness first (short timeouts), then timing against the unified kernel
mkdir -p gpurun_out
echo "== golden + C1 tests"; timeout 150 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "c1 or golden or overlapped_back or fm" 2>&1 | tail -4
[ ${PIPESTATUS[0]} -ne 0 ] && { echo "WS kernel not correct: stopping"; exit 0; }
echo "== ws (default)";      timeout 60 python scripts/microbench.py dec chainfm 2>&1 | tail -2
echo "== ws again";          timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
echo "== unified (ORION_B200_NO_WS)"; ORION_B200_NO_WS=1 timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
V=$PWD/orion-sdr_b200/variants
for v in $(ls $V 2>/dev/null | sed 's/liborion_b200_//; s/\.so//'); do echo "== $v"; ORION_B200_LIB=$V/liborion_b200_$v.so timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1; done
for ns in 8 7 6; do echo "== ws stages $ns"; ORION_B200_STAGES=$ns timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1; done
echo "== isolated"; OVERLAP=0 timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -4 | tee gpurun_out/ws_pytest.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 120 $CMD > gpurun_out/plain_ws.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_ws -s 30 -c 1 -f -o gpurun_out/r02_ws $CMD > gpurun_out/ncu_ws.log 2>&1
echo "ncu exit=$?"; bash scripts/ncu_export.sh r02_ws; ls -la gpurun_out | tail -8
