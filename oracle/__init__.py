"""CPU oracle for the bullet-js hot path - TEST INFRASTRUCTURE ONLY.

Nothing under `bullet_js_b200/` imports this package.  Allowed importers:
`tests/`, `__graft_entry__.smoke()`, and bench.py's cpu_baseline / --impl reference legs.
"""
