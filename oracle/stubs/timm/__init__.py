"""Stand-in for the `timm` package (absent from this image): the reference's Swin files import three helpers from
timm.models.layers only (swinunet_icl.py:9, vision_transformer.py:26)."""
