"""Utilities: timers, environment flags, checkpoint helpers."""
