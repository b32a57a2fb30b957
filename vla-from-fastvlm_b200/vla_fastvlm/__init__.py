"""B200-native FastVLA policy forward behind the VLA-from-FastVLM plugin surface."""
