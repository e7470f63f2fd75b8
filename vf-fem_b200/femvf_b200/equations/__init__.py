"""Equation helpers (mirror of ``femvf.equations``)."""
