"""Drop-in mirrors of the reference's models/ package (same class names, constructor arguments, forward
signatures and state_dict keys), backed by the b200caps CUDA kernels."""
