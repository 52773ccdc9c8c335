import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        km = d.get('roofline_context', {}).get('kernel_ms') or d['roofline'].get('kernel_ms')
        print('n', d['n_gpus'], 'value %.2fM' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], 'e2e %.2fM' % (d['e2e']['value']/1e6), 'frac', round(d['roofline']['frac'] or 0, 3), {k: round(v, 4) for k, v in km.items()})
    else:
        print(l.rstrip()[:200])
