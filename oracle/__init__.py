"""CPU oracle of the reference hot path — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this package; boosting_rcnn_b200/ never does.
"""
