"""ORACLE — test infrastructure only (see oracle/disco_oracle.hpp). Never imported by sasktran2_b200."""
