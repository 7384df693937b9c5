/* Prints the derived RADAU5 constants exactly as the reference computes them at run time
 * (thirdparty/radau5/radau5_io.c:347-370, glibc sqrt/pow) so that they can be pasted as constexpr doubles
 * into assimulo_b200/csrc/aens_radau5.cuh.   gcc -O0 tools/radau_consts.c -lm && ./a.out */
#include <math.h>
#include <stdio.h>
int main(void) {
    volatile double six = 6., e81 = 81., nine = 9., third = .33333333333333331, three = 3.;
    double sq6 = sqrt(six), p1 = pow(e81, third), p2 = pow(nine, third);
    double c1 = (4. - sq6) / 10., c2 = (sq6 + 4.) / 10.;
    double alph = (12. - p1 + p2) / 60., beta = (p1 + p2) * sqrt(three) / 60., cno;
    printf("c1=%.17g c2=%.17g c1mc2=%.17g c1m1=%.17g c2m1=%.17g\n", c1, c2, c1 - c2, c1 - 1, c2 - 1);
    printf("dd1=%.17g dd2=%.17g dd3=%.17g\n", -(sq6 * 7. + 13.) / 3., (sq6 * 7. - 13.) / 3., -.33333333333333331);
    printf("u1=%.17g ", 1. / ((p1 + 6. - p2) / 30.));
    cno = alph * alph + beta * beta;
    alph /= cno; beta /= cno;
    printf("alph=%.17g beta=%.17g\n", alph, beta);
    return 0;
}
