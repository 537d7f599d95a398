"""Mirror of the reference's code/utils package for the ICL hot path (losses)."""
