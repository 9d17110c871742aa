"""Runs every `-m gpu` test in its own process with a timeout, so one trapped / hung kernel cannot
poison the CUDA context of the others.  Prints a one-line verdict per test and a summary; writes
gpurun_out/gpu_tests.json.     python tools/run_gpu_tests.py [pytest-file ...] [-k expr]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    kexpr = None
    if "-k" in args:
        i = args.index("-k")
        kexpr = args[i + 1]
        args = args[:i] + args[i + 2:]
    files = args or ["tests"]
    cmd = [sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + files
    if kexpr:
        cmd += ["-k", kexpr]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True).stdout
    ids = []
    for l in out.splitlines():          # one process per test FUNCTION (all its parametrisations)
        if "::" in l:
            fn = l.strip().split("[")[0]
            if fn not in ids:
                ids.append(fn)
    results = {}
    t_all = time.time()
    for nid in ids:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", nid, "-rf", "--no-header", "-p", "no:cacheprovider"],
                               cwd=ROOT, capture_output=True, text=True, timeout=600)
            ok = p.returncode == 0
            tail = "" if ok else "\n".join((p.stdout + p.stderr).splitlines()[-40:])
        except subprocess.TimeoutExpired:
            ok, tail = False, "TIMEOUT (600 s)"
        results[nid] = {"ok": ok, "s": round(time.time() - t0, 1), "tail": tail}
        print(("PASS " if ok else "FAIL ") + nid + f"  ({results[nid]['s']} s)", flush=True)
        if not ok:
            print("    " + tail.replace("\n", "\n    "), flush=True)
    npass = sum(r["ok"] for r in results.values())
    print(f"== {npass}/{len(results)} gpu tests passed in {time.time() - t_all:.0f} s")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "gpu_tests.json"), "w"), indent=1)
    sys.exit(0 if npass == len(results) else 1)


if __name__ == "__main__":
    main()
