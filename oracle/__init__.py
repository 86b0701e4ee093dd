"""CPU oracle for the feature-extraction hot path.  TEST INFRASTRUCTURE ONLY -- see DESIGN.md.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs only.
"""
