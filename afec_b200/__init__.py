"""afec_b200: B200-native low-level audio descriptor engine (AFEC `--level low` hot path)."""
