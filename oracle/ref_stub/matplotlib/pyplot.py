"""Empty stand-in, see the package docstring."""
