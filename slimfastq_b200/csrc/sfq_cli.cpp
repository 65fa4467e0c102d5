// placeholder main; replaced below
int main() { return 0; }
