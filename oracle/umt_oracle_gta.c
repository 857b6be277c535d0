/* umt_oracle_gta.c — TEST INFRASTRUCTURE (see umt_oracle.c header).
 * Grey-transport-acceleration pieces of the oracle; filled in below. */
int orc_gta_placeholder(void) { return 0; }
