"""CPU oracle for the CP tensor-regression fit iteration.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tensor_regression_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker or as the timed CPU baseline — never as the product path.

Parity pinning: ``tr_oracle`` is checked (tests/test_oracle_pinned.py) against
  * golden vectors produced by running the UNMODIFIED reference modules
    (/root/reference/standard_tensor_regression.py,
    /root/reference/multinomial_tensor_regression.py) through the tensorly
    stand-in in ``tensorly_standin`` (script: oracle/make_golden.py, fixtures:
    tests/golden/*.npz), and
  * the saved known-answer log of the reference's demo_TensorRegression.ipynb
    cell 8 (final L-BFGS loss 0.041904340578888165).
"""
