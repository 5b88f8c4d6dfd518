"""CPU oracle for the bullet-js hot path - TEST INFRASTRUCTURE ONLY.

Nothing under `bullet_js_b200/` imports this package.  Allowed importers:
`tests/` (including `tests/golden/make_golden.py`, which writes the fixtures the tests read, and
`tests/check_shard_gpu.py`), `__graft_entry__.smoke()` and bench.py's cpu_baseline / --impl reference legs.

  js_literal.py / jsvalue.py   statement-level restatement over JS-like Python values
  bullet_oracle.c / typed.py   the same on the typed struct-of-arrays format (C, built by oracle/Makefile)
  minijs/ + ref_runner.py      an ECMAScript-subset interpreter that runs /root/reference/src/*.js
                               unmodified; the source of tests/golden/*.json.gz
"""
