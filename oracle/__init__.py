"""CPU oracle for the ILRMA / AuxIVA iterative demixing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ssspy_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
(or as the CPU arm that is timed *beside* the product), never as the product.

What it is: an fp64 NumPy restatement, written from the update equations, of
the reference path named in SURVEY.md section 8(a):

* ``oracle.linalg``           <- ssspy/linalg/{_solve,inv,eigh,lqpqm}.py, ssspy/special/psd.py
* ``oracle.spatial``          <- ssspy/bss/_update_spatial_model.py (IP1, IP2, ISS1, ISS2, IPA)
* ``oracle.projection_back``  <- ssspy/algorithm/{projection_back,minimal_distortion_principle}.py
* ``oracle.ilrma``            <- ssspy/bss/ilrma.py (Gauss / Student-t / GGD ILRMA, MM/ME, partitioning, all spatial modes)
* ``oracle.iva``              <- ssspy/bss/iva.py  (AuxLaplaceIVA / AuxGaussIVA)
* ``oracle.mnmf``             <- ssspy/bss/mnmf.py (FastGaussMNMF, IP1/IP2 diagonaliser)
* ``oracle.fdica``            <- ssspy/bss/fdica.py (AuxLaplaceFDICA) and
                                 ssspy/algorithm/permutation_alignment.py (correlation-based solver)

Unlike the reference it never materialises the (I,N,N,N,J) broadcast
temporaries (ssspy/bss/ilrma.py:1500-1505); contractions are einsum/matmul.

Parity pinning: the reference ships no offline golden vectors for this path
(its ``target.npz`` regression files are network downloads).
The oracle is therefore pinned against outputs of the reference itself, run in
the build container by ``tests/golden/make_golden.py`` and ``make_golden_fdica.py`` (scripts committed, vectors
committed as ``tests/golden/*.npz``) and by the docstring known-answer values of
``inv2`` / ``eigh2`` (ssspy/linalg/inv.py:20-37, ssspy/linalg/eigh.py:53-74,131-152).
``tests/test_oracle_golden.py`` checks every fixture.
"""

from . import fdica, ilrma, iva, linalg, mnmf, projection_back, spatial  # noqa: F401

EPS = 1e-10
