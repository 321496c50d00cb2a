import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print(sys.argv[1], 'ms/cycle', round(d['ms_per_step'],4))
        for k in d['kernels']:
            if 'halo' in k['kernel'] or 'allgather' in k['kernel'] or '@L0' in k['kernel']:
                print("   %-28s x%.0f %8.4f ms"%(k['kernel'],k['launches_per_cycle'],k['ms']))
