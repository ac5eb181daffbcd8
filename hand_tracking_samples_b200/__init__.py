"""B200-native replacement for the CNN hot path of IntelRealSense/hand_tracking_samples."""
