"""CPU oracle for the genz_tokenize hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from genz_tokenize_b200 (the product).
"""
