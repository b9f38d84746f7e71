"""eoe_b200 -- B200-native scoring / loss / AUC hot path of liznerski/eoe behind the reference's ADTrainer hooks.

Host code is Python/PyTorch (memory, streams, torch.distributed); all compute is hand-written sm_100a CUDA in
eoe_b200/libeoe_b200.so, reached through the C ABI of include/eoe_b200.h.  No CPU fallback exists.
"""
__version__ = "0.1.0"
