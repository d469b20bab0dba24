/* TEST INFRASTRUCTURE ONLY: command-line front end of the host-emulated aln pipeline (libaln_emul.so). */
int pansvr_fc_aln_main(int argc, char **argv);
int main(int argc, char **argv) { return pansvr_fc_aln_main(argc, argv); }
