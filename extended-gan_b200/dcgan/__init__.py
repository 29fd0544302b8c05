"""Drop-in mirror of the reference's ``dcgan`` package for the hot path only: the three conv nets."""
