for so in flux_b200/lib/variants/lib_*.so; do echo $so; FLUXB200_LIB=$PWD/$so timeout 600 python tools/bench_configs.py ${AB_CONFIGS:-c3} 2>&1 | grep -v linear | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['config'], {k:round(v,2) for k,v in d.items() if k in ('Mrays_per_s','Msamples_per_s','nodes_per_segment','prim_tests_per_segment','set_scene_s')})
"; done
