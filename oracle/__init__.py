"""CPU oracle (NumPy/SciPy restatement of the reference layer math).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs -- never by the product
package ``gcn_fmri_decoding_b200``.  PARITY UNPINNED at the TensorFlow boundary;
see ``oracle/layers_np.py`` for what is and is not pinned.
"""
