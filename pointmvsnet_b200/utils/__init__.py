"""Mirrors of the reference utils the PointFlow path touches: feature fetcher, kNN, file formats."""
