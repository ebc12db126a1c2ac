"""bench.py's bookkeeping that needs no GPU: the hash that ties an ncu capture (profiles/roofline_traffic.json) to the
kernel sources it was taken at ignores comments and white space, differs per kernel, and the committed records are
well formed."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_source_hash_ignores_comments_and_white_space(tmp_path, monkeypatch):
    pkg = tmp_path / "mpilattice-boltzmann_b200" / "csrc"
    pkg.mkdir(parents=True)
    src = os.path.join(ROOT, "mpilattice-boltzmann_b200", "csrc")
    for name in ("lbm_cell.cuh", "lbm_kernels.cuh", "lbm_stepsk.cuh"):
        shutil.copy(os.path.join(src, name), pkg / name)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    before = {k: bench.kernel_source_hash(k) for k in (2, 5, 7)}
    assert before[2] == before[5] != before[7]                  # kernel 7 adds lbm_stepsk.cuh
    p = pkg / "lbm_stepsk.cuh"
    text = p.read_text()
    p.write_text("// a new comment\n/* and a block\n comment */\n" + text.replace("\n", "\n  \n", 3))
    assert {k: bench.kernel_source_hash(k) for k in (2, 5, 7)} == before
    p.write_text(text.replace("constexpr int kHalo = 4;", "constexpr int kHalo = 5;"))
    after = {k: bench.kernel_source_hash(k) for k in (2, 5, 7)}
    assert after[7] != before[7] and after[2] == before[2]      # a code change disowns kernel 7's capture only


def test_committed_traffic_records_are_well_formed():
    top = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    for kernel, rec in top["kernels"].items():
        steps = rec["algorithmic_bytes_per_launch"] / (72 * rec["nx"] * rec["ny"])
        assert steps in (1, 2, 3, 4), kernel
        per_cell_step = rec["dram_bytes_per_launch"] / (rec["nx"] * rec["ny"]) / steps
        assert abs(per_cell_step - rec["dram_bytes_per_cell_per_step"]) < 1e-6
        # a kernel that fuses k steps per pass moves about 72 / k bytes per cell and step (+ redundant columns / rows)
        assert 0.99 * 72 / steps <= per_cell_step < 1.08 * 72 / steps + 0.5, (kernel, per_cell_step)
        assert len(rec["source_hash"]) == 16 and rec["capture"].startswith("profiles/")
        assert os.path.exists(os.path.join(ROOT, rec["capture"].split(" ")[0]))
