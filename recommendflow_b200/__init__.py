"""recommendflow_b200 -- B200-native feature-to-embedding hot path behind RecommendFlow's layer API."""
