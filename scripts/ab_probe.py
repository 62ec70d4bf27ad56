"""A/B timing of a GEMM env switch in ONE gpurun call: alternates child processes (same box, same thermal state)."""
import json, os, subprocess, sys
env_name, values = sys.argv[1], sys.argv[2].split(",")
sizes = sys.argv[3:] or ["4096x4096x4096"]
res = {v: [] for v in values}
for rep in range(3):
    for v in values:
        env = dict(os.environ, **{env_name: v})
        p = subprocess.run([sys.executable, "scripts/gemm_probe.py", "child", "cg2_bn128", "0", *sizes], capture_output=True, text=True, env=env, timeout=300)
        for line in p.stdout.splitlines():
            if line.startswith("{"):
                d = json.loads(line)
                res[v].append((d["M"], d.get("ms"), d.get("max_rel")))
for v in values:
    print(env_name, "=", v, [(m, round(ms, 4) if ms else None) for m, ms, _ in res[v]], "err", res[v][0][2])
