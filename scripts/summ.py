import json,sys,os
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f"{os.path.basename(f):24s} value {d['value']:9.1f} Gcmp/s  ms/step {d['ms_per_step']:8.4f}  kern_ms {r.get('kernel_ms',0):.4f} frac {r.get('frac') and round(r['frac'],3)} e2e {d['e2e']['value']:.1f} var {d['config'].get('variant')} launches {d.get('gpu_launches')} clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']} matched {d.get('matched_per_step')} n_gpus {d['n_gpus']}")
        if d.get('cpu_baseline'): print('   cpu', d['cpu_baseline'])
    except Exception as e:
        print(f, 'ERR', e)
