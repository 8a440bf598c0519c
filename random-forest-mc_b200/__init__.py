"""B200-native engine for the RandomForestMC fit / score / vote hot path (drop-in Python API)."""

__version__ = "1.4.0"  # wire-format version of the reference this engine mirrors
__engine_version__ = "0.1.0"
