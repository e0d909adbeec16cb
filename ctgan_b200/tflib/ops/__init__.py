"""Op surface of the reference's tflib.ops package (TG/tflib/ops/)."""
