"""vidchapters_b200 — B200-native Vid2Seq train step (drop-in for model.vid2seq.Vid2Seq)."""
