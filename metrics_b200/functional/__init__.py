"""Stateless functional front-ends (reference: src/torchmetrics/functional/)."""
