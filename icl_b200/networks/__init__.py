"""Mirror of the reference's code/networks package for the ICL hot path (3D U-Net family)."""
