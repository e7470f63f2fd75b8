"""Residual definitions (mirror of ``femvf.residuals``)."""
