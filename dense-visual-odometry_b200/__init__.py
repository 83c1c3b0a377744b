"""B200-native photometric-alignment hot path behind the reference's estimator API."""
__all__ = []
