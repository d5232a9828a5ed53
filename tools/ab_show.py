import json, sys
for line in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/ab_bench.log'):
    line = line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        d = json.loads(line)
        print('  ', {k: (round(v['reg_per_s_device']), round(v['ms_per_step_device'], 3)) for k, v in d.items() if isinstance(v, dict) and 'reg_per_s_device' in v},
              {k: round(v, 3) for k, v in d['profile_ms'].items()}, d['iters'], d['launches_per_step'])
    elif line: print(line[:300])
