"""Run GPU test groups in separate processes with timeouts (a trapped kernel poisons its CUDA context; isolation keeps
the other groups meaningful) and collect the logs under gpurun_out/.  Usage: python tools/gpu_diag.py [group ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
GROUPS = {
    "elem": ["tests/test_kernels_gpu.py", "-k", "bn_ or avgpool or head or flat or weight_prep"],
    "conv_fwd": ["tests/test_kernels_gpu.py", "-k", "conv_forward"],
    "conv_dgrad": ["tests/test_kernels_gpu.py", "-k", "conv_dgrad"],
    "conv_wgrad": ["tests/test_kernels_gpu.py", "-k", "conv_wgrad"],
    "stem": ["tests/test_kernels_gpu.py", "-k", "stem"],
    "engine": ["tests/test_engine_gpu.py"],
    "dropin": ["tests/test_dropin_gpu.py"],
}


def main():
    os.makedirs(OUT, exist_ok=True)
    groups = sys.argv[1:] or list(GROUPS)
    summary = {}
    for g in groups:
        t0 = time.time()
        cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "--timeout", "600", "--durations=8", "-p", "no:cacheprovider"] + GROUPS[g]
        try:
            r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
            out, code = r.stdout + r.stderr, r.returncode
        except subprocess.TimeoutExpired as e:
            out, code = (e.stdout or b"").decode(errors="replace") + "\nTIMEOUT", -9
        with open(os.path.join(OUT, f"diag_{g}.log"), "w") as f:
            f.write(out)
        tail = out.strip().splitlines()[-1] if out.strip() else ""
        summary[g] = dict(code=code, seconds=round(time.time() - t0, 1), tail=tail)
        print(g, summary[g], flush=True)
    with open(os.path.join(OUT, "diag_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
