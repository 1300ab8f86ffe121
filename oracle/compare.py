"""Test infrastructure (like everything under oracle/): comparison of token sequences under the north-star tolerance.
Ids must be identical except where the oracle's decision margin at the first differing step is inside the tolerance
(bf16 operands flip near-ties).  After an exempted flip the rest of that row is not comparable
(the two decoders follow different prefixes) and is skipped.  Used by tests/, __graft_entry__.smoke() and bench.py's
result check only."""
import torch


def compare_greedy(seq, ref_seq, ref_margins, tol):
    """Returns (n_exact_rows, n_exempt_rows, failures[list of (row, step, margin)])."""
    exact = exempt = 0
    failures = []
    for r in range(ref_seq.size(0)):
        diff = (seq[r] != ref_seq[r]).nonzero()
        if diff.numel() == 0:
            exact += 1
            continue
        t = int(diff[0])
        m = float(ref_margins[r, t])
        if m < tol:
            exempt += 1
        else:
            failures.append((r, t, m))
    return exact, exempt, failures


def compare_beam(seq, ref_seq, ref_rel_margins, tol):
    """Beam search: a row must equal the oracle's unless the oracle's smallest relative decision margin along that image's
    search (oracle.sample_beam(..., return_margins=True)) is inside the tolerance.
    Returns (n_exact_rows, n_exempt_rows, failures[list of (row, margin)])."""
    exact = exempt = 0
    failures = []
    for r in range(ref_seq.size(0)):
        if torch.equal(seq[r], ref_seq[r]):
            exact += 1
        elif float(ref_rel_margins[r]) < tol:
            exempt += 1
        else:
            failures.append((r, float(ref_rel_margins[r])))
    return exact, exempt, failures


