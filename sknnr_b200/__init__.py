"""sknnr_b200: B200-native query-time hot path of sknnr."""
