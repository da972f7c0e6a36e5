"""m3l_b200 — B200-native (sm_100a) implementation of M3L's VTMAE/VTT train step."""
__version__ = "0.1.0"
