"""Drop-in mirrors of the reference's ``dmi/model`` modules (projector, hypernet, lora, mmmodel)."""
