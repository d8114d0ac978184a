"""Layer wrappers with the reference parameter names (Conv1d + BN, SharedMLP), forward over the C ABI."""
