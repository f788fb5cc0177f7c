"""DESIGN.md / INTEGRATION.md stay in step with the sources: every runtime switch the library reads is documented, no
documented switch is dead (unless DESIGN.md says it is gone), and the entry-point count quoted matches the header."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GONE = {"JP_APPLY_GRAPH", "JP_FUSED_DOT", "JP_SPMM_RUNS_DB", "JP_SPMM_RUNS_INTERIOR", "JP_SPMM_TILE", "JP_TEST_UNVALIDATED"}  # named as removed
NOT_SWITCHES = {"JP_CUDA_ERROR", "JP_OK", "JP_INVALID_ARGUMENT", "JP_INVALID_STATE", "JP_NCCL_ERROR"}


def _switches_in_sources():
    found = set()
    for path in glob.glob(os.path.join(ROOT, "juliapetra.jl_b200", "csrc", "*.c*")) + glob.glob(os.path.join(ROOT, "juliapetra.jl_b200", "csrc", "*.h")):
        src = open(path).read()
        found |= set(re.findall(r'(?:getenv|env_int)\("(JP_[A-Z0-9_]+)"', src))
    for path in glob.glob(os.path.join(ROOT, "juliapetra.jl_b200", "*.py")):
        found |= set(re.findall(r'environ\.get\("(JP_[A-Z0-9_]+)"', open(path).read()))
    return found


def test_every_runtime_switch_is_documented_and_none_is_dead():
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    documented = set(re.findall(r"JP_[A-Z0-9_]+", design)) - NOT_SWITCHES
    used = _switches_in_sources()
    assert used <= documented, f"undocumented switches: {sorted(used - documented)}"
    assert documented - used <= GONE, f"documented but not read anywhere: {sorted(documented - used - GONE)}"
    assert not (GONE & used), f"switches DESIGN.md calls removed are still read: {sorted(GONE & used)}"


def test_entry_point_count_quoted_in_design_matches_the_header():
    header = open(os.path.join(ROOT, "include", "jpetra_b200.h")).read()
    n = len(re.findall(r"^int32_t jp_[a-z0-9_]+\(", header, flags=re.M))
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    m = re.search(r"jpetra_b200\.h` \((\d+) entry points\)", design)
    assert m and int(m.group(1)) == n, (m and m.group(1), n)


def test_profiles_named_in_design_exist():
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    names = set(re.findall(r"`(?:profiles/)?((?:r[12]_|ncu_)[A-Za-z0-9_.{},*]+\.(?:json|txt|log|csv))`", design))
    missing = []
    for name in names:
        if any(ch in name for ch in "{}*"):
            continue  # a pattern, not a file
        if not os.path.exists(os.path.join(ROOT, "profiles", name)):
            missing.append(name)
    assert not missing, missing
