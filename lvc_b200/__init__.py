"""lvc_b200: Blackwell-native pseudo-label mining hot path of prannaykaul/lvc."""
__version__ = "0.2.0"
