"""CPU oracle of the koala_b200 signal path -- TEST INFRASTRUCTURE ONLY (see koala_oracle.c header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from .oracle import Oracle, OracleBatch, OracleModel, build_oracle, oracle_lib_path  # noqa: F401
