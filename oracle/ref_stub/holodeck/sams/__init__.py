"""Stub ``holodeck.sams`` package: only hosts the compiled reference ``sam_cyutils``."""
