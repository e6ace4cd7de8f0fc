#!/usr/bin/env python
"""Drop-in proof, step 3: run the reference's own test programs -- compiled UNMODIFIED by integration/Makefile
against integration/_build/libccv_dropin.so (the reference's L1..L5 + this backend as the 8th one, SM100 the only GPU
backend) -- on a GPU box and tabulate PASS / SKIP / FAIL / CRASH per test case.  Each program runs once; its main
(integration/case_fork_main.c) isolates crashes in forked children and prints `@@ <STATUS> <case name>` after each case's own
output, which is kept as the detail of a FAIL / CRASH (the REQUIRE_... message with both values, or the failed assertion).

  python integration/run_reference_tests.py [--only int.cudnn,int.cublas] [--match "half precision"] [--budget-s 600]
                                            [--retry-from gpurun_out/dropin_reference_tests.json]   # only what failed there
  python integration/run_reference_tests.py --manifest /root/reference    # build container: case names per program -> _build/cases.json
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
MARK = re.compile(r"^@@ (PASS|SKIP|FAIL|CRASH) (.*?)(?: \((?:signal|exit|outside)[^)]*\))?\s*$")


def manifest(ref):
    out = {}
    for kind in ("int", "unit"):
        d = os.path.join(ref, "test", kind, "nnc")
        for f in sorted(os.listdir(d)):
            if f.endswith(".tests.c"):
                out["%s.%s" % (kind, f[:-2])] = re.findall(r'^TEST_CASE\("(.*?)"\)', open(os.path.join(d, f)).read(), re.M)
    os.makedirs(BUILD, exist_ok=True)
    json.dump(out, open(os.path.join(BUILD, "cases.json"), "w"), indent=0)
    print("wrote", os.path.join(BUILD, "cases.json"), sum(len(v) for v in out.values()), "cases")


def parse(text):
    res, detail, buf = {}, {}, []
    for raw in text.replace("\r", "\n").replace("\b", "").split("\n"):
        m = MARK.match(raw.strip())
        if not m:
            if raw.strip():
                buf.append(raw.rstrip())
            continue
        res[m.group(2)] = m.group(1)
        if m.group(1) in ("FAIL", "CRASH"):
            detail[m.group(2)] = (raw.strip() + " | " + " / ".join(buf[-4:]))[-700:]
        buf = []
    return res, detail


def _omp_threads():
    """threads for the reference's CPU side of a test case (its CPU_REF comparison runs): the CPUs this process may use -- affinity cut by
    a cgroup quota -- and at most 16: a container on a 128-thread host otherwise starts 128 spinning threads per parallel region"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return str(max(1, min(16, n)))


def run(cmd, timeout):
    t0 = time.time()
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", _omp_threads())
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, cwd=BUILD, env=env)
        return p.stdout.decode(errors="replace"), time.time() - t0
    except subprocess.TimeoutExpired as e:
        return (e.stdout or b"").decode(errors="replace") + "\n[program time limit]", time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--manifest", metavar="REFERENCE_ROOT")
    ap.add_argument("--only", default="")
    ap.add_argument("--match", default="", help="substring filter on case names (argv[1] of the test programs)")
    ap.add_argument("--retry-from", default="", help="a previous report: run only the cases that failed / crashed / did not run there")
    ap.add_argument("--budget-s", type=float, default=600.0)
    ap.add_argument("--case-timeout", type=int, default=120)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "gpurun_out", "dropin_reference_tests.json"))
    args = ap.parse_args()
    if args.manifest:
        return manifest(args.manifest)
    cases = json.load(open(os.path.join(BUILD, "cases.json")))
    progs = sorted(f for f in os.listdir(BUILD) if f.endswith(".tests"))
    if args.only:
        keep = args.only.split(",")
        progs = [p for k in keep for p in progs if k in p]  # in the order asked for
    retry = None
    if args.retry_from:
        prev = json.load(open(args.retry_from))["programs"]
        retry = {p: sorted(set(r.get("failed", [])) | set(r.get("crashed", {})) | set(r.get("not_run_cases", []))) for p, r in prev.items()}
        progs = [p for p in progs if retry.get(p)]
    t_start = time.time()
    report = {}
    for prog in progs:
        names = cases.get(prog, [])
        runs = [[os.path.join(BUILD, prog), args.match, str(args.case_timeout)]]
        if retry is not None:
            runs = [[os.path.join(BUILD, prog), n, str(args.case_timeout)] for n in retry[prog]]
            names = retry[prog]
        res, detail, secs = {}, {}, 0.0
        for cmd in runs:
            left = args.budget_s - (time.time() - t_start)
            if left <= 0:
                break
            text, dt = run(cmd, left)
            r, d = parse(text)
            secs += dt
            for k, v in r.items():
                if retry is None or k in names:
                    res[k] = v
            detail.update({k: v for k, v in d.items() if k in res})
        tally = {k: sum(1 for v in res.values() if v == k) for k in ("PASS", "SKIP", "FAIL", "CRASH")}
        missing = [n for n in names if n not in res and (not args.match or args.match in n)]
        tally["not_run"] = len(missing)
        report[prog] = {"tally": tally, "seconds": round(secs, 1), "passed": sorted(k for k, v in res.items() if v == "PASS"), "skipped": sorted(k for k, v in res.items() if v == "SKIP"),
                        "failed": sorted(k for k, v in res.items() if v == "FAIL"), "crashed": {k: detail.get(k, "") for k, v in res.items() if v == "CRASH"},
                        "fail_detail": {k: detail.get(k, "") for k, v in res.items() if v == "FAIL"}, "not_run_cases": missing}
        print("%-28s %s  (%.1fs)" % (prog, tally, secs), flush=True)
    total = {k: sum(r.get("tally", {}).get(k, 0) for r in report.values()) for k in ("PASS", "SKIP", "FAIL", "CRASH", "not_run")}
    print("TOTAL", total)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"total": total, "programs": report}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
