timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -4
timeout 600 python scripts/perf_configs.py cfg2 cfg1 2>&1 | python scripts/_pp_perf.py | grep -v pdl0 | grep -v "cfg2_search2\|cfg2_single" 
echo "=== NO PAIR"; KEDS_NO_PAIR=1 timeout 600 python scripts/perf_configs.py cfg1 2>&1 | python scripts/_pp_perf.py | grep -v pdl0 | head -3
