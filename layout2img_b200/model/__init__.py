"""Drop-in module surfaces of the hot path (same class names, constructor arguments, forward
signatures and state_dict keys as the reference's model/ package)."""
