import json, sys
d = json.load(open('/root/repo/gpurun_out/.last_call.json'))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
print(d['status'], 'rc', d['rc'], 'run_s', d.get('run_s'), 'left', d.get('gpu_minutes_left'), 'cmd', d.get('cmd', '')[:80])
print(d['stdout_tail'][-n:])
if d['stderr_tail'].strip():
    print('--- stderr'); print(d['stderr_tail'][-1200:])
