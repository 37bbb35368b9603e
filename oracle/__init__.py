"""CPU oracle for the DReg-NeRF registration hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline.  The product path
(``dreg-nerf_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Parity status (also recorded in DESIGN.md):
  * R1-R3, R5-R10 (``oracle/regtr.py``): PINNED.  The restatement is checked
    against the reference's own modules imported from /root/reference
    (``oracle/ref_shim.py``) by ``oracle/make_goldens.py``; the outputs of the
    *reference* are committed under ``tests/golden/``.
  * R4 (MinkowskiEngine voxel average), A1-A5 (tiny-cuda-nn hash grid / MLPs,
    nerfacc 0.3.5 ray marching): PARITY UNPINNED.  Those algorithms live in
    third-party CUDA wheels that are neither vendored in /root/reference nor
    installable here; the restatements follow the published algorithms and the
    reference's call sites.  The only in-repo golden vector (transmittance
    docstring, conerf/utils/nerfacc_utils.py:56-63) is checked.
  * ``oracle/extract_c.c`` (``make -C oracle`` -> ``oracle/_c/``): the A1 / A5
    arithmetic once more in C with OpenMP (fp32, no early outs), tied to the
    Python oracles' fixtures by ``tests/test_oracle_golden.py``; used for the
    full-size 128^3 parity test and as bench.py's timed CPU marcher.
"""
