"""Test-infrastructure package: CPU oracle of the reference hot path (see oracle.py). Never imported by the product."""
