"""mchammer_b200 (placeholder during bring-up)."""
