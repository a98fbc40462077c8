"""Empty: the reference never calls a transform on the hot path."""
