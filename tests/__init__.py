"""CPU suite (-m "not gpu") and GPU parity suite (-m gpu) of the PointFlow path."""
