import sys, json
for line in sys.stdin:
    if line.startswith('{"metric'):
        d = json.loads(line)
        if d["steps"] == 2: continue
        print('BENCH', d['config']['symbols_per_gpu'], '%.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'],
              'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], 'cpu', d['cpu_baseline'] and '%.4g' % d['cpu_baseline']['value'])
    else:
        print(line, end='')
