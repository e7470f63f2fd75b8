"""Model definitions (mirror of ``femvf.models``)."""
