"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of DDMAL/text_alignment's alignment path (textSeqCompare.py:13-177) used as
the checker for the CUDA implementation.  Nothing under ``text_alignment_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.

Parity status: PINNED -- against golden vectors produced by the unmodified reference
(``tests/golden/make_golden.py``) and against the live reference when ``/root/reference``
exists (``oracle.ref_loader``).
"""
