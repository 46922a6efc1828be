"""CPU oracle for the frame -> KL-f8 latent -> binary-code hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product package (``symbols-from-video_b200``) never imports this package and
raises if its CUDA library is missing.

What it restates
----------------
The reference (matt-suncy/symbols-from-video) is pure Python; the arithmetic of
the path lives in a third-party dependency that is NOT under /root/reference:
PyTorch (``torch==2.3.0+cu118`` pinned at reference ``requirements.txt:172``;
this container has torch 2.11.0).  The oracle therefore restates the reference's
*call sequence* as state-dict-driven functional code on ``torch.nn.functional``
CPU fp32 ops (``kl_f8.py``, ``rbvae.py``, ``frames.py``) and, for the primitive
ops themselves (conv2d / group_norm / softmax-attention / LSTM cell), carries an
independent explicit-loop numpy restatement of their published definitions
(``primitives_np.py``) that the tests check the torch ops against on small cases.

Parity pinning
--------------
The reference has no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned the second way the task allows:
against outputs of the *unmodified reference modules* imported in the build
container from /root/reference (``ref_shim.py``), with the vectors committed
under ``tests/golden/`` together with the generating script
(``oracle/make_golden.py``).  ``tests/test_oracle_golden.py`` re-checks the
oracle against those vectors without needing /root/reference.

The reference itself also travels: ``build_ref.py`` (run by ``__graft_entry__.build()`` in the build container)
copies the 8 hot-path reference files and the sample video byte for byte into the git-ignored ``oracle/_ref/``;
``ref_shim.py`` imports the unmodified classes from there when /root/reference is absent (GPU box), which is what
``bench.py --impl reference`` / its ``cpu_baseline`` leg time (``kind: "reference"``) next to the port.
"""
