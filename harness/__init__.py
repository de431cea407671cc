"""Measurement harnesses for BASELINE configs 3-5 (decoder layer, full QAT step,
13B sizing).  They restate just enough of the reference's LLaMA model to put the
hot path under a realistic caller; they are not part of the product package."""
