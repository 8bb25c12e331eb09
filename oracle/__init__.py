"""CPU oracle for the OptimalControlMPS hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy restatement of the
reference's Bose-Hubbard tDMRG + optimal-control path (reference files
``src/BH_tDMRG.cpp``, ``src/OptimalControl.cpp``, ``src/ControlBasis.cpp``,
``include/ControlBasisFactory.hpp``, ``include/SeedGenerator.hpp``,
``include/BH_sites.h``) together with the ITensor v2 semantics those files rely
on.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``optimalcontrolmps_b200`` never does.

Pinning status
--------------
* forward evolution, cost, fidelities: pinned to the reference's own golden
  vectors ``tests/CostTests.cpp:75-197`` (L=5) -- see ``tests/test_oracle_goldens.py``.
* ControlBasis / chopped-sine basis: pinned to ``tests/ControlBasisTests.cpp`` goldens.
* gradient / Hessian: pinned only self-consistently (finite differences at the
  reference's own tolerances, ``tests/GradientTests.cpp``, ``tests/HessianTests.cpp``).
* chi-capped truncation, centre-move rank drops and the ``exactApplyMPO`` compression
  follow ITensor v2 as recalled in SURVEY.md appendix A; ITensor is a third-party
  dependency that is not vendored, not version-pinned and not installable here, so
  for those details **parity is unpinned** beyond the goldens above.
"""
