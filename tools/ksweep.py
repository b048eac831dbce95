import sys, json, subprocess
for k in (8, 16, 20, 32, 64, 128):
    p = subprocess.run([sys.executable, "bench.py", "--steps", str(k), "--warmup", "5", "--no-cpu-baseline"], capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(k, round(d["value"], 1), round(d["ms_per_step"], 4), [round(x, 2) for x in d["repetitions_ms"]], round(d["e2e"]["value"], 1), flush=True)
    except Exception as e:
        print(k, "ERR", p.stderr[-500:])
