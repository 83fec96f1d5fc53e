"""In-situ timeline of one pass (bf_run) at the bench workload: BF_TRACE_KERNELS=1 makes the library record an event
after every launch; this prints microseconds per launch next to the api.cu line that issued it.

    python tools/trace_pass.py [n_profiles] [max_dist] [key=value options ...]
"""
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if os.environ.get("BF_TRACE_KERNELS") != "1":
    env = dict(os.environ, BF_TRACE_KERNELS="1")
    res = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    src = (ROOT / "breakfast_b200" / "csrc" / "api.cu").read_text().splitlines()
    sys.stdout.write(res.stdout)
    blocks = res.stderr.split("[bf trace] total")
    last = blocks[-2] if len(blocks) >= 2 else res.stderr
    for line in last.splitlines():
        m = re.match(r"\[bf trace\] api.cu:(\d+)\s+([\d.]+) us", line)
        if m:
            ln = int(m.group(1))
            text = " ".join(src[ln - 3:ln]).strip()
            k = re.findall(r"(k_\w+|cudaLaunchCooperativeKernel|launch_\w+)", text)
            print(f"{float(m.group(2)):9.1f} us  api.cu:{ln:<5d} {k[-1] if k else text[:60]}")
    tail = [l for l in res.stderr.splitlines() if "total" in l]
    print(tail[-1] if tail else res.stderr[-2000:])
    sys.exit(res.returncode)

from breakfast_b200 import _native, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1
opts = {k: int(v) for k, v in (a.split("=") for a in sys.argv[3:])}
indptr, indices, n_cols = synth.generate(n, seed=1).csr()
with _native.Context(**opts) as ctx:
    ctx.upload_csr(indptr, indices, n_cols)
    for _ in range(4):
        st = ctx.run_sync(d)
    print(f"n={n} d={d} opts={opts}: ms_total {st.ms_total:.4f} sort {st.ms_sort:.4f} pack {st.ms_pack:.4f} pairs {st.ms_pairs:.4f} "
          f"verify {st.ms_verify:.4f} cc {st.ms_cc:.4f}  edges {st.n_edges} components {st.n_components}")
