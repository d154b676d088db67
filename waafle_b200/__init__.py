"""waafle_b200: B200-native scoring and clade-assignment engine for WAAFLE's orgscorer path."""

__version__ = "0.1.0"
