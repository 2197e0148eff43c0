"""Sampling estimators (reference nerfacc/estimators/)."""
