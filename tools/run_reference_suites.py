#!/usr/bin/env python
"""
Runs the reference's OWN test modules on the B200 backend (INDIGO_TEST_BACKENDS=b200,
backends/__init__.py:9-14) and prints / stores the pass counts.

    python tools/run_reference_suites.py [--stride K] [--out profiles/r02_reference_suites.json]

The reference comes from /root/reference or, on the GPU box, from oracle/_ref/reference_pkg.zip
(see tests/golden/refshim.py).  Test infrastructure: the product never imports any of this.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
sys.path.insert(0, REPO)

SUITES = ["indigo/backends/test_backends.py", "indigo/test_operators.py", "indigo/test_transforms.py"]


def run_suite(root, rel, stride=1, offset=0, timeout=3000):
    env = dict(os.environ, INDIGO_TEST_BACKENDS=os.environ.get("IB200_REFSUITE_BACKENDS", "b200"), IB200_REFSUITE_STRIDE=str(stride),
               IB200_REFSUITE_OFFSET=str(offset),
               PYTHONPATH=os.pathsep.join([os.path.join(REPO, "tests"), REPO, os.environ.get("PYTHONPATH", "")]))
    cmd = [sys.executable, "-m", "pytest", "-p", "refsuite_plugin", "-q", "-x", "--no-header", "-p", "no:cacheprovider",
           "--rootdir", root, "-W", "ignore", os.path.join(root, rel)]
    if stride == 1:
        cmd.remove("-x")
    t0 = time.time()
    p = subprocess.run(cmd, cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    tail = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|skipped|xfailed|xpassed|deselected|errors?)", tail)}
    fails = [l for l in p.stdout.splitlines() if l.startswith(("FAILED", "ERROR"))][:20]
    return dict(suite=rel, rc=p.returncode, seconds=round(time.time() - t0, 1), summary=tail, counts=counts, failures=fails,
                output_tail=p.stdout[-3000:] if p.returncode else "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import refshim
    root = refshim.reference_root()
    if root is None:
        raise SystemExit("reference not available")
    res = [run_suite(root, rel, args.stride) for rel in SUITES]
    for r in res:
        print("%-40s rc=%d %6.1fs  %s" % (r["suite"], r["rc"], r["seconds"], r["summary"]))
        for f in r["failures"]:
            print("   ", f)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(dict(backend="b200", stride=args.stride, results=res), open(args.out, "w"), indent=1)
    sys.exit(max(r["rc"] for r in res))


if __name__ == "__main__":
    main()
