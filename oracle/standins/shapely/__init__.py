"""Stand-in for shapely 1.6.4: ring coordinates and min/max bounds only."""
