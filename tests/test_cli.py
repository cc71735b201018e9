"""CLI runner (u_pol_b200/cli.py): the reference's --mol/--dir flags and log.out wire format
(polarization.py:207-226, :261-263) on a force field + PDB written by the test itself."""
import os

import numpy as np
import pytest

from u_pol_b200 import synth
from u_pol_b200.topology import build_system

XML = """<ForceField>
 <AtomTypes>
  <Type name="w-O" class="OW" element="O" mass="0"/>
  <Type name="w-H" class="HW" element="H" mass="0"/>
  <Type name="w-M" class="MW" mass="0"/>
  <Type name="w-OD" class="OWD" mass="0"/>
 </AtomTypes>
 <Residues>
  <Residue name="HOH">
   <Atom name="O" type="w-O"/><Atom name="H1" type="w-H"/><Atom name="H2" type="w-H"/>
   <Atom name="M" type="w-M"/><Atom name="OD" type="w-OD"/>
   <VirtualSite type="average3" index="3" atom1="0" atom2="1" atom3="2" weight1="{w1}" weight2="{w2}" weight3="{w2}"/>
   <Bond from="0" to="1"/><Bond from="0" to="2"/>
  </Residue>
 </Residues>
 <NonbondedForce coulomb14scale="0" lj14scale="0">
  <Atom type="w-O" charge="{qo}" sigma="1" epsilon="0"/>
  <Atom type="w-H" charge="{qh}" sigma="1" epsilon="0"/>
  <Atom type="w-M" charge="{qm}" sigma="1" epsilon="0"/>
  <Atom type="w-OD" charge="{qd}" sigma="1" epsilon="0"/>
 </NonbondedForce>
 <DrudeForce>
  <Particle type1="w-OD" type2="w-O" charge="{qd}" polarizability="{alpha}" thole="1.3"/>
 </DrudeForce>
</ForceField>
"""


def _write_inputs(root):
    m = synth.load_monomer("water")
    os.makedirs(os.path.join(root, "water"))
    with open(os.path.join(root, "water", "water.xml"), "w") as fh:
        fh.write(XML.format(w1=0.589781071, w2=0.2051094645, qo=m["q_core"][0], qh=m["q_core"][1], qm=m["q_core"][3],
                            qd=m["q_shell"][0], alpha=m["alpha"][0]))
    s = synth.water_box(2, seed=7, density=20.0)
    with open(os.path.join(root, "water", "water.pdb"), "w") as fh:
        n = 0
        for mol in range(2):
            for a, name in enumerate(("O", "H1", "H2")):
                n += 1
                x, y, z = s.r_core[mol, a] * 10.0
                fh.write(f"HETATM{n:5d}  {name:<3s} HOH {'AB'[mol]}{mol + 1:4d}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00\n")
        fh.write("END\n")


def test_reader_on_generated_inputs(tmp_path):
    _write_inputs(str(tmp_path))
    s = build_system(str(tmp_path / "water" / "water.pdb"), str(tmp_path / "water" / "water.xml"))
    assert s.r_core.shape == (2, 4, 3) and s.site_names == ["O", "H1", "H2", "M"]
    assert np.count_nonzero(s.q_shell) == 2 and not s.has_screened_pairs


@pytest.mark.gpu
def test_cli_log_wire_format(tmp_path):
    from u_pol_b200 import cli
    from u_pol_b200.engine import UindSystem

    _write_inputs(str(tmp_path))
    log = tmp_path / "log.out"
    u = cli.main(["--mol", "water", "--dir", str(tmp_path), "--log", str(log)])
    lines = open(log).read().splitlines()
    assert lines[1] == "%%%%%%%%%%% STARTING WATER U_IND CALCULATION %%%%%%%%%%%%"
    py = [l for l in lines if l.startswith("Python U_ind = ")]
    assert py == [f"Python U_ind = {u:.7f} kJ/mol"]            # polarization.py:262
    s = build_system(str(tmp_path / "water" / "water.pdb"), str(tmp_path / "water" / "water.xml"), positions_dtype="float32")
    _, info = UindSystem.from_drude_system(s).minimize(tol=1e-8)
    assert u == pytest.approx(info["U_ind"], rel=1e-8) and u < 0
