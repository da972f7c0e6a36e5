"""TEST INFRASTRUCTURE — empty shim so /root/reference/models/pretrain_models.py imports
(stable-baselines3 is not installed here; only base-class names are needed at import time)."""
