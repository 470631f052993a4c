"""Host-side mirror of the reference's `anime_3dkenburns` package for the hot path."""
