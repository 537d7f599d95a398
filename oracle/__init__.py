"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the ICL hot path (zhuye98/ICL, reference at /root/reference) used as
the parity checker for the CUDA product in ``icl_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it.  Nothing under ``icl_b200/`` imports or calls anything here; the product path
fails loudly when its CUDA extension is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures (SURVEY.md §4,
§8c).  The restatement in ``oracle/restate.py`` is therefore pinned by *executing the
unmodified reference* in the build container (``oracle/ref_import.py``) on deterministic
synthetic parameters/inputs (``oracle/synth.py``) and committing its outputs as fixtures
under ``tests/golden/`` (generator: ``oracle/make_golden.py``).  ``tests/test_oracle_*``
check the restatement against those fixtures on CPU, and — when /root/reference is present
— against the live reference.

All floating-point arithmetic of the reference lives in a third-party dependency:
PyTorch (reference pin torch==1.9.1+cu111, README.md:20; this image: torch 2.11.0 CPU ops)
plus three MONAI 1.0.1 symbols (Conv factory, DropPath, ensure_tuple_rep — restated in
``oracle/stubs``) and MedPy 0.4.0's ``metric.binary.dc`` (restated in ``restate.dice_metric``).
"""
