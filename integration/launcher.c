/* main() of the drop-in binary: the reference's own main (main.c:95, renamed at compile time
 * with -Dmain=tagdust_main when the reference is built as a library). */
int tagdust_main(int argc, char* argv[]);
int main(int argc, char* argv[]) { return tagdust_main(argc, argv); }
