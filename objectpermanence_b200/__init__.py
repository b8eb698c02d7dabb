"""objectpermanence_b200 -- B200-native OPNet temporal-reasoning hot path."""
__version__ = "0.1.0"
