"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything under ``oracle/``.  The product
package (``nessai_b200``) never does and fails loudly without its CUDA library.

Contents
--------
``shims/``      stand-ins for packages the reference needs but this image lacks:
                ``glasflow`` (+ ``glasflow.nflows``: a PyTorch restatement of
                the nflows arithmetic the reference delegates to, NOT in
                /root/reference -- **parity unpinned** at that boundary),
                and import stubs for matplotlib / seaborn / cycler.
``refenv.py``   puts ``baseline/_ref`` (the UNMODIFIED reference, pip-installed
                with ``--no-deps --target``) and the shims on ``sys.path``.
``flow_numpy.py`` independent float64 numpy restatement of the folded
                eval-mode flow used to cross-check the shim and the kernels.
"""
