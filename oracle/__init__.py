"""TEST INFRASTRUCTURE -- CPU oracle for the attention hot path (see network_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (metal-flash-attention_b200/) never does.
"""
from .oracle import *  # noqa: F401,F403
