"""lmfit stand-in: import-time placeholders (golden generation only)."""
