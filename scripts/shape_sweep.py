"""Randomised shape sweep of the public model API against the oracle (tests/shape_cases.py): prints one line per case;
exit code 1 if any fails.  usage: shape_sweep.py [n_cases] [seed] [only_case | -1] [variants | big]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import shape_cases

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
only = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) >= 0 else None
big = len(sys.argv) > 4 and sys.argv[4] == "big"               # benchmark-sized shapes (slow: the oracle runs on the CPU)
variants = len(sys.argv) > 4 and sys.argv[4] == "variants"     # optional paths (dropout, scheduled sampling, SCST, use_bn, diverse beam)
fails = 0
for case, c in enumerate(shape_cases.cases(n_cases, seed, big)):
    if only is not None and case != only:
        continue
    tag = f"case {case}: " + " ".join(f"{k}={v}" for k, v in c.items())
    try:
        if variants:
            v = shape_cases.VARIANTS[case % len(shape_cases.VARIANTS)]
            tag += " variant=" + v
            msg, summary = shape_cases.run_variant(case, c, v)
        else:
            msg, summary = shape_cases.run_case(case, c)
        if msg:
            fails += 1
            print("FAIL", tag, "|", "; ".join(msg), flush=True)
        else:
            print("ok  ", tag, "|", summary, flush=True)
    except Exception as e:
        fails += 1
        print("EXC ", tag, "|", type(e).__name__, str(e)[:300], flush=True)
        torch.cuda.synchronize()
print("failures:", fails)
sys.exit(1 if fails else 0)
