"""CPU oracle for the RISER read-classification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``riser_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or the timed CPU baseline -- never as the product path.

It is a numpy / torch-CPU restatement of the reference algorithm
(comprna/riser: ``riser/preprocess.py``, ``riser/model.py``,
``riser/nets/cnn.py``, ``riser/control.py``); every function cites the
reference ``file:line`` it follows.

Pinning status: the reference ships NO tests, golden vectors or trained
weights for this path, so by the reference's own
material the parity is "unpinned".  The oracle is pinned instead against
outputs of the reference itself, imported read-only in the build container:
``tests/golden/make_golden.py`` runs the real ``preprocess.SignalProcessor``,
``model.Model``, ``nets.cnn.ConvNet`` and a verbatim driver of
``control.py:31-93`` on seeded synthetic inputs and commits the results as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this package
against those files bit-for-bit (preprocessing) / to 1e-6 (network).
"""
