"""print the key fields of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        e = d.get("e2e") or {}
        r = d.get("roofline") or {}
        print(f"{path}: n_gpus={d.get('n_gpus')} ms/step={d.get('ms_per_step'):.3f} value={d.get('value'):.4g} "
              f"e2e_ms={e.get('ms_per_step')} e2e={e.get('value'):.4g} frac={r.get('frac')} "
              f"launches={d.get('gpu_launches')} clocks={d.get('clocks')} phases={d.get('phases_ms')}")
        if d.get("cpu_baseline"):
            print("   cpu_baseline:", d["cpu_baseline"])
