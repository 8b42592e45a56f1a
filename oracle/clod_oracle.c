/* placeholder, filled in below */
int clod_oracle_version(void) { return 1; }
