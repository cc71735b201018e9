"""Batched scan driver (u_pol_b200/scan.py): parser on CPU, the full pipeline on the GPU."""
import os

import numpy as np
import pytest

from u_pol_b200 import scan, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
EXC = os.path.join(GOLD, "psi4_excerpts")


def test_parser_reproduces_the_reference_generated_geometries():
    """generated_pdbs/2mer-0+<id>.pdb (written by the reference's wrapper.py) == our parse + %8.3f."""
    z = np.load(os.path.join(GOLD, "dimers312.npz"))
    for did, e_sapt in ((10, -2.43673156), (102, None), (389, None)):
        coords, e_ind = scan.parse_sapt_outfile(os.path.join(EXC, f"2mer-0+{did}.out"))
        assert coords.shape == (18, 3) and e_ind is not None
        if e_sapt is not None:
            assert e_ind == e_sapt
        i = int(np.nonzero(z["dimer_id"] == did)[0][0])
        assert e_ind == pytest.approx(float(z["sapt_induction"][i]), rel=1e-12)
        nm = scan.pdb_rounded_nm(coords, positions="float64").reshape(2, 9, 3)
        assert np.abs(nm - z["r_core"][i]).max() < 1e-12


@pytest.mark.gpu
def test_scan_reproduces_reference_csv_rows(tmp_path):
    z = np.load(os.path.join(GOLD, "dimers312.npz"))
    out = tmp_path / "induction_data.csv"
    ids, sapt, md = scan.run_scan(EXC, synth.load_monomer("imidazole"), outfile=str(out), tol=1e-6)
    assert ids.tolist() == [10, 102, 389]
    for did, e_md in zip(ids, md):
        i = int(np.nonzero(z["dimer_id"] == did)[0][0])
        assert abs(e_md - z["md_induction"][i]) <= 5e-4 * abs(z["md_induction"][i])
    rows = np.loadtxt(out, delimiter=",", skiprows=1)
    assert open(out).readline().strip() == "SAPT_Induction,MD_Induction,DimerID"
    assert rows.shape == (3, 3) and np.array_equal(rows[:, 2], [10, 102, 389])
    assert np.allclose(rows[:, 1], md, rtol=0, atol=0)
