"""TEST INFRASTRUCTURE ONLY: CPU oracle of the Salience-DETR encoder hot path (see oracle/oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.
"""
