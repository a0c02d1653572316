"""CPU oracle for the MIPSFusion per-frame neural-field hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mipsfusion_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the CPU baseline -- never as the thing shipped.

Each function restates (in plain torch fp32 on the CPU, or numpy for integer
work) what the reference computes, citing the reference file:line it follows.
Citations are relative to the reference checkout (``/root/reference``).

Parity status
-------------
The reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4).  The oracle is therefore pinned as follows:

* everything that is the reference's *own* Python (decoder, rendering, losses,
  samplers, RandomOptimizer, the Mesher weight blend) is checked against the
  reference modules imported unchanged from ``/root/reference`` through the
  shims in ``oracle/shims`` -- see ``tests/golden/make_golden.py``, which wrote
  the committed fixtures in ``tests/golden/`` from the reference's outputs;
* the HashGrid / Frequency encodings live in the un-vendored third-party
  dependency ``tinycudann==1.7`` (reference ``environment.yaml:74``).  Its
  published algorithm is restated in ``oracle/hashgrid.py`` (torch) and,
  independently, in ``oracle/hashgrid_ref.c`` (plain C, real uint32
  arithmetic); the two are cross-checked, and the known-answer level table
  of SURVEY.md Appendix A is asserted.  There is no tcnn binary or tcnn golden
  vector offline, so for those two encodings: **parity unpinned**.
"""
