"""Model pieces: the MoE layer, gates, experts and auxiliary losses."""
