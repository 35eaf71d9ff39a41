"""astropy stand-in (golden generation only)."""
