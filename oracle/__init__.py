"""TEST INFRASTRUCTURE: CPU oracles for the embedding hot path.

* ``oracle.port``  – this repo's C++ restatement of the reference algorithm
  (``hetu_port.cc`` -> ``libhetu_port.so``), with numpy restatements of the
  reference CUDA kernels in ``oracle.ops_port``.
* ``oracle.ref``   – loader for the reference's own hetu_cache + server handler,
  compiled from /root/reference by ``oracle/Makefile`` into ``oracle/_ref``.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``herald_b200``) never does.
"""
