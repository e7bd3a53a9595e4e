/* placeholder translation unit until the v1 restatement lands */
int futbol_v1_oracle_present(void) { return 0; }
