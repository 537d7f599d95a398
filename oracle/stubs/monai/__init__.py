"""Test-only stand-in for the three MONAI 1.0.1 symbols the reference ICL heads import.
Lives under oracle/: test infrastructure, never imported by the product package."""
