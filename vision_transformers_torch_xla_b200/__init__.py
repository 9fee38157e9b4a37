"""B200-native ViT training hot path."""
