"""CPU oracle (NumPy/SciPy restatement of the reference layer math).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs -- never by the product
package ``gcn_fmri_decoding_b200``.  PARITY UNPINNED at the TensorFlow boundary;
see ``oracle/layers_np.py`` for what is and is not pinned.

* ``layers_np.py``   -- the oracle: forward, backward, loss, whole training step (fp32 "as run" and fp64).
* ``ref_loader.py``  -- build container only: imports the reference's host modules (``graph``, ``coarsening``) and
  compiles pieces of its TensorFlow-bound source by AST (``cgcnn`` on a stand-in for ``tf``, ``base_model.fit`` /
  ``predict`` / ``evaluate``, ``model_perf``, ``BestCheckpointSaver``, the builders of ``model.py``).
* ``tf_shim.py``     -- the NumPy and torch stand-ins for the ~25 TensorFlow ops the reference's layer code calls.
* ``make_golden*.py`` -- generators of ``tests/golden/`` (graphs / recursion / layer vectors, experiment configurations,
  vectors of the reference source on the stand-ins).
"""
