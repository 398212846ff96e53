"""Checker-side experiments (they execute oracle/, so they live under tests/, not in the product or
its tools): trajectory parity of the sampled and the full-decode path, loader measurement against
the reference's loader.  Run as scripts: python tests/experiments/<name>.py"""
