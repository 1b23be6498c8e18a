"""Host-side engine around the C-ABI kernels: flat-parameter optimiser, device actor pools,
rollout buffers, learners (PyTorch holds memory/streams; the arithmetic is in libparl_b200.so)."""
