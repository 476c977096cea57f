"""Mapping objects (georeferenced images) backed by device-resident coordinate planes."""
