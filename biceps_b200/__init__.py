"""biceps_b200 -- B200-native BICePs posterior-sampling hot path behind the reference's API."""
__version__ = "0.1.0"
name = "biceps_b200"
