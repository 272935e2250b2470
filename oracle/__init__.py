"""CPU oracle for the QuatMpc / ConvexMpc GRF solve.  TEST INFRASTRUCTURE ONLY — see altro_ref.h."""
