import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "b-spline-two-e_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_terminal_summary(terminalreporter):
    """strict per-entry error of every CSR comparison of the run (parity_utils.PARITY_STATS)"""
    import json
    from parity_utils import PARITY_STATS, REL_TOL
    if not PARITY_STATS:
        return
    n = sum(q["entries"] for q in PARITY_STATS)
    over = sum(q["strict_over_1e-12"] for q in PARITY_STATS)
    worst = max(PARITY_STATS, key=lambda q: q["strict_max"])
    tr = terminalreporter
    tr.write_line(f"parity: {len(PARITY_STATS)} CSR comparisons, {n} entries; scaled max "
                  f"{max(q['scaled_max'] for q in PARITY_STATS):.2e}; strict per-entry max {worst['strict_max']:.2e} "
                  f"({worst['what']}); entries over {REL_TOL:.0e} strict: {over} ({over / max(n, 1):.2e} of all)")
    path = os.environ.get("BS2E_PARITY_REPORT")
    if path:
        agg = {}
        for q in PARITY_STATS:
            key = q["what"].split(" ")[0] if q["what"] else "?"
            a = agg.setdefault(key, {"comparisons": 0, "entries": 0, "scaled_max": 0.0, "strict_max": 0.0, "strict_over_1e-12": 0})
            a["comparisons"] += 1
            a["entries"] += q["entries"]
            a["scaled_max"] = max(a["scaled_max"], q["scaled_max"])
            a["strict_max"] = max(a["strict_max"], q["strict_max"])
            a["strict_over_1e-12"] += q["strict_over_1e-12"]
        with open(path, "w") as f:
            json.dump({"total_entries": n, "strict_over_1e-12": over, "worst": worst, "by_label": agg}, f, indent=1)


# Small deterministic basis_input namelists (the namelist is the seed; the
# reference has no random inputs).  Chosen to hit the edge cases of the path:
# radial truncation of the second electron (r_2_max), the large-r angular
# hole (r_all_l / max_l2), odd and even L, z_pol on and off, full on and off.
SMALL_CASES = {
    "tiny_k4": dict(k=4, m=2, Z=1, h_max=1.0, r_max=6.0, k_GL=8, max_k=2, max_L=1, max_l_1p=1,
                    max_l2=1, CAP_eta=0j, CAP_r_0=4.0, full=True, z_pol=True),
    "trunc_k5": dict(k=5, m=2, Z=2, h_max=1.0, r_max=8.0, k_GL=9, max_k=3, max_L=2, max_l_1p=2,
                     max_l2=1, r_2_max=5.0, r_all_l=6.0, CAP_eta=5e-3 + 0j, CAP_r_0=5.0,
                     full=False, z_pol=False),
    "wide_k6": dict(k=6, m=3, Z=2, h_max=0.8, r_max=10.0, k_GL=12, max_k=5, max_L=3, max_l_1p=3,
                    max_l2=3, r_2_max=6.0, CAP_eta=1e-3 + 0j, CAP_r_0=7.0, full=False, z_pol=True),
    # spline order above the BASELINE configs: more than 32 n_c slots per site and more than 256
    # candidate columns (several candidate passes, chunked warp scans)
    "order10": dict(k=10, m=2, Z=1, h_max=1.0, r_max=9.0, k_GL=13, max_k=2, max_L=1, max_l_1p=1, max_l2=1,
                    CAP_eta=1e-3 + 0j, CAP_r_0=6.0, full=False, z_pol=True),
    # s waves only, one multipole: the smallest angular structure (one (l1,l2) group, one symmetry)
    "swave_k0": dict(k=5, m=2, Z=2, h_max=1.0, r_max=7.0, k_GL=8, max_k=0, max_L=0, max_l_1p=0, max_l2=0,
                     CAP_eta=0j, CAP_r_0=5.0, full=False, z_pol=True),
    # many multipoles on a tiny basis: the 21- and 31-register instantiations of the site kernel
    # (max_k of BASELINE configs 4 and 5), with k_GL >= k + max_k/2 so that the inner rule is exact
    "multipoles20_k5": dict(k=5, m=2, Z=1, h_max=1.0, r_max=6.0, k_GL=15, max_k=20, max_L=2, max_l_1p=3,
                            max_l2=3, r_2_max=4.0, CAP_eta=2e-3 + 0j, CAP_r_0=4.0, full=False, z_pol=True),
    "multipoles30_k4": dict(k=4, m=2, Z=2, h_max=1.0, r_max=5.0, k_GL=19, max_k=30, max_L=1, max_l_1p=2,
                            max_l2=2, CAP_eta=0j, CAP_r_0=4.0, full=True, z_pol=True),
}


@pytest.fixture(scope="session")
def small_cases():
    return SMALL_CASES
