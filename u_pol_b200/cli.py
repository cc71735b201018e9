"""Command-line runner with the reference's flags and log wire format (SURVEY 8f "next" #4).

Mirrors ``main()`` of ``U_pol/polarization.py:194-266``: ``--mol {water,acnit,imidazole,...} --dir``,
builds the arrays (``util`` equivalents over the OpenMM-free reader), minimises, evaluates ``Uind`` and
appends to ``log.out`` the lines the reference writes (``:261-263``).  The OpenMM SCF cross-check
(``util.U_ind_omm``, needs OpenMM) is printed as "n/a" when OpenMM is not importable.

    python -m u_pol_b200.cli --mol acnit --dir /path/to/benchmarks/OpenMM
"""
from __future__ import annotations

import argparse
import logging
import os
import time
from datetime import datetime

import numpy as np


def main(argv=None):
    from . import polarization as pol
    from . import util

    ap = argparse.ArgumentParser(description="Calculate U_ind = U_self + U_es for a selected molecule.")
    ap.add_argument("--mol", type=str, required=True,
                    choices=["water", "acnit", "imidazole", "imidazole2", "imidazole3", "pyrazine"])
    ap.add_argument("--dir", type=str, default="../benchmarks/OpenMM")
    ap.add_argument("--log", type=str, default="log.out")
    args = ap.parse_args(argv)

    logger = logging.getLogger("u_pol_b200.cli")
    logger.setLevel(logging.INFO)
    for h in list(logger.handlers):
        logger.removeHandler(h)
    fh = logging.FileHandler(args.log)
    fh.setFormatter(logging.Formatter("%(message)s"))
    logger.addHandler(fh)
    logger.info(f"Log started at: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}")
    pol.set_constants()
    mol = args.mol
    logger.info(f"%%%%%%%%%%% STARTING {mol.upper()} U_IND CALCULATION %%%%%%%%%%%%")
    logger.info("-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=")
    system = util.build_system(os.path.join(args.dir, mol, mol + ".pdb"), os.path.join(args.dir, mol, mol + ".xml"),
                               positions_dtype="float32")
    Rij, Dij = util.get_Rij_Dij(system)
    Qi_core, Qi_shell, Qj_core, Qj_shell = util.get_QiQj(system)
    k, u_scale = util.get_pol_params(system)
    start = time.time()
    Dopt = pol.drudeOpt(Rij, np.ravel(Dij), Qi_shell, Qj_shell, Qi_core, Qj_core, u_scale, k)
    U_ind = float(pol.Uind(Rij, Dopt, Qi_shell, Qj_shell, Qi_core, Qj_core, u_scale, k))
    logger.info(f"U_POL_B200 minimiser Time: {time.time() - start:.3f} seconds.")
    logger.info("OpenMM U_ind = n/a (OpenMM not used)")
    logger.info(f"Python U_ind = {U_ind:.7f} kJ/mol")
    logger.info("=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=\n")
    fh.close()
    print(f"U_ind = {U_ind:.7f} kJ/mol")
    return U_ind


if __name__ == "__main__":
    main()
