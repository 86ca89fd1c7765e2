"""Drop-in mirror of the reference's `opencood` registry surface for the hot path."""
