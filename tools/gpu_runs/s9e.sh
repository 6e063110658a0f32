# overlapped SW schedule (RRTMGX_SW_SPLIT=2): parity, then timing against the fused default and the plain split path
timeout 600 python -m pytest tests/test_sw_gpu.py -q -x -k "split" > gpurun_out/s9e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9e_tests.log
tail -4 gpurun_out/s9e_tests.log
timeout 600 python tools/sweep.py 65536 72 "fused:" "split1:RRTMGX_SW_SPLIT=1" "ovl_b3:RRTMGX_SW_SPLIT=2,RRTMGX_SW_DOWN_BLOCKS=3" "ovl_b2:RRTMGX_SW_SPLIT=2,RRTMGX_SW_DOWN_BLOCKS=2" "ovl_b4:RRTMGX_SW_SPLIT=2,RRTMGX_SW_DOWN_BLOCKS=4" "ovl_b6:RRTMGX_SW_SPLIT=2,RRTMGX_SW_DOWN_BLOCKS=6" "ovl_b3_u0:RRTMGX_SW_SPLIT=2,RRTMGX_SW_DOWN_BLOCKS=3,RRTMGX_SW_UP=0" "ovl_b3_u2:RRTMGX_SW_SPLIT=2,RRTMGX_SW_DOWN_BLOCKS=3,RRTMGX_SW_UP=2" "fused_again:" > gpurun_out/s9e_sweep.jsonl 2> gpurun_out/s9e_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/s9e_sweep.jsonl"):
    if l.startswith("{"):
        r=json.loads(l); print(f"{r['cfg']:<12} both_ms={r['both_ms']:.2f} lw_ms={r['lw_ms']:.2f} sw_ms={r['sw_ms']:.2f} status={r['status']}")
PY
tail -3 gpurun_out/s9e_sweep.err
