"""CPU oracle for the FEAR-XS per-frame inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``feartracker_b200`` (the product) may import
this package.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as
the checker / the reported CPU baseline, never as the thing shipped.

Contents
--------
``fbnet_c``      restatement of the third-party ``mobile_cv`` FBNet-V2 ``fbnet_c``
                 backbone (facebookresearch/mobile-vision @ 51804a6873ae, pinned by
                 the reference's ``requirements.txt:8``; NOT vendored in the reference).
``fear_oracle``  self-contained torch-CPU restatement of the reference's
                 ``FEARNet`` / ``FEARBoxCoder`` / ``FEARTracker`` arithmetic, driven by
                 a plain state_dict.  It travels to the GPU box (``/root/reference``
                 does not).
``ref_shims``    ``sys.modules`` stubs that let the reference's OWN source import
                 unchanged from ``/root/reference`` in the build container.
``make_golden``  runs the real reference through the shims, checks the restatement
                 against it bit-for-bit and writes ``tests/golden/*``.

Parity pinning status
---------------------
The reference ships no tests or golden vectors for this path.
The restatement is therefore pinned against *outputs of the reference's own source
run in the build container* (``make_golden.py``; results committed under
``tests/golden/``).  The one piece that cannot be pinned by running reference code is
the ``mobile_cv`` backbone itself ("parity unpinned" for that dependency): its
structure is fixed by the checkpoint's parameter names/shapes (strict load) and by
the op graph of the CoreML models the reference ships; BN eps = 1e-5 is assumed.
"""
