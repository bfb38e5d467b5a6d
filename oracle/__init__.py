"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/cfd_oracle.hpp). Never imported by cfd_demo_b200."""
