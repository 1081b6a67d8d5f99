/* ORACLE (test infrastructure only -- see suews_oracle.h).
 *
 * The two MINPACK drivers AnOHM uses, restated from the reference's own copy of MINPACK
 * (SUEWS-SourceCode/suews_util_minpack.f95, J. Burkardt's Fortran 90 rendering of More / Garbow / Hillstrom):
 *   lmdif1 (:3477-3616) -> lmdif (:2942-3476), fdjac2 (:740-860), qrfac (:4743-4936), lmpar (:3617-3920),
 *                          qrsolv (:4937-5154), enorm (:381-429)
 *   hybrd1 (:1413-1557) -> hybrd (:862-1411), fdjac1 (:560-739, dense branch: hybrd1 sets ml = mu = n-1), qform
 *                          (:4642-4742), dogleg (:171-380), r1updt (:5266-5499), r1mpyq (:5155-5265)
 * Same statements in the same order; array expressions (SUM, DOT_PRODUCT, whole-array assignments) are the sequential
 * loops gfortran emits without -ffast-math.  The user functions of AnOHM never return a negative iflag and nprint is 0,
 * so those exits are not restated.  Sizes: n <= SO_MP_MAXN unknowns, m <= SO_MP_MAXM residuals.
 */
#include "suews_oracle.h"
#include <math.h>
#include <stddef.h>
/* epsilon(), tiny(), huge() of REAL(8) */
#define SO_EPSMCH 2.220446049250313e-16
#define SO_DWARF 2.2250738585072014e-308
#define SO_GIANT 1.7976931348623157e308

#define A_(a, lda, i, j) ((a)[(i) + (size_t)(lda) * (j)])

/* minpack.f95:381-429: enorm = sqrt(sum(x(1:n)**2)) */
static double enorm(int n, const double *x)
{
   double s = 0;
   int i;
   for (i = 0; i < n; i++) s = s + x[i] * x[i];
   return sqrt(s);
}

static double dotp(int n, const double *a, const double *b)
{
   double s = 0;
   int i;
   for (i = 0; i < n; i++) s = s + a[i] * b[i];
   return s;
}

/* minpack.f95:4743-4936 */
static void qrfac(int m, int n, double *a, int lda, int pivot, int *ipvt, double *rdiag, double *acnorm)
{
   const double epsmch = SO_EPSMCH;
   double wa[SO_MP_MAXN], r8_temp[SO_MP_MAXM], ajnorm, temp;
   int i, j, k, kmax, minmn, i4_temp;
   for (j = 0; j < n; j++) acnorm[j] = enorm(m, &A_(a, lda, 0, j));
   for (j = 0; j < n; j++) { rdiag[j] = acnorm[j]; wa[j] = acnorm[j]; }
   if (pivot)
      for (j = 0; j < n; j++) ipvt[j] = j;
   minmn = m < n ? m : n;
   for (j = 0; j < minmn; j++) {
      if (pivot) {
         kmax = j;
         for (k = j; k < n; k++)
            if (rdiag[kmax] < rdiag[k]) kmax = k;
         if (kmax != j) {
            for (i = 0; i < m; i++) r8_temp[i] = A_(a, lda, i, j);
            for (i = 0; i < m; i++) A_(a, lda, i, j) = A_(a, lda, i, kmax);
            for (i = 0; i < m; i++) A_(a, lda, i, kmax) = r8_temp[i];
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            i4_temp = ipvt[j];
            ipvt[j] = ipvt[kmax];
            ipvt[kmax] = i4_temp;
         }
      }
      ajnorm = enorm(m - j, &A_(a, lda, j, j));
      if (ajnorm != 0.0) {
         if (A_(a, lda, j, j) < 0.0) ajnorm = -ajnorm;
         for (i = j; i < m; i++) A_(a, lda, i, j) = A_(a, lda, i, j) / ajnorm;
         A_(a, lda, j, j) = A_(a, lda, j, j) + 1.0;
         for (k = j + 1; k < n; k++) {
            temp = dotp(m - j, &A_(a, lda, j, j), &A_(a, lda, j, k)) / A_(a, lda, j, j);
            for (i = j; i < m; i++) A_(a, lda, i, k) = A_(a, lda, i, k) - temp * A_(a, lda, i, j);
            if (pivot && rdiag[k] != 0.0) {
               double t2, q;
               temp = A_(a, lda, j, k) / rdiag[k];
               t2 = 1.0 - temp * temp;
               rdiag[k] = rdiag[k] * sqrt(t2 > 0.0 ? t2 : 0.0);   /* max(0, 1 - temp**2) */
               q = rdiag[k] / wa[k];
               if (0.05 * (q * q) <= epsmch) {
                  rdiag[k] = enorm(m - j - 1, &A_(a, lda, j + 1, k));
                  wa[k] = rdiag[k];
               }
            }
         }
      }
      rdiag[j] = -ajnorm;
   }
}

/* minpack.f95:4937-5154 */
static void qrsolv(int n, double *r, int ldr, const int *ipvt, const double *diag, const double *qtb, double *x,
                   double *sdiag)
{
   double wa[SO_MP_MAXN], c, s, cotan, t, qtbpj, temp, sum2;
   int i, j, k, l, nsing;
   for (j = 0; j < n; j++) {
      for (i = j; i < n; i++) A_(r, ldr, i, j) = A_(r, ldr, j, i);
      x[j] = A_(r, ldr, j, j);
   }
   for (j = 0; j < n; j++) wa[j] = qtb[j];
   for (j = 0; j < n; j++) {
      l = ipvt[j];
      if (diag[l] != 0.0) {
         for (k = j; k < n; k++) sdiag[k] = 0.0;
         sdiag[j] = diag[l];
         qtbpj = 0.0;
         for (k = j; k < n; k++) {
            if (sdiag[k] != 0.0) {
               if (fabs(A_(r, ldr, k, k)) < fabs(sdiag[k])) {
                  cotan = A_(r, ldr, k, k) / sdiag[k];
                  s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                  c = s * cotan;
               } else {
                  t = sdiag[k] / A_(r, ldr, k, k);
                  c = 0.5 / sqrt(0.25 + 0.25 * (t * t));
                  s = c * t;
               }
               A_(r, ldr, k, k) = c * A_(r, ldr, k, k) + s * sdiag[k];
               temp = c * wa[k] + s * qtbpj;
               qtbpj = -s * wa[k] + c * qtbpj;
               wa[k] = temp;
               for (i = k + 1; i < n; i++) {
                  temp = c * A_(r, ldr, i, k) + s * sdiag[i];
                  sdiag[i] = -s * A_(r, ldr, i, k) + c * sdiag[i];
                  A_(r, ldr, i, k) = temp;
               }
            }
         }
      }
      sdiag[j] = A_(r, ldr, j, j);
      A_(r, ldr, j, j) = x[j];
   }
   nsing = n;
   for (j = 0; j < n; j++) {
      if (sdiag[j] == 0.0 && nsing == n) nsing = j;
      if (nsing < n) wa[j] = 0.0;
   }
   for (j = nsing - 1; j >= 0; j--) {
      sum2 = dotp(nsing - j - 1, &wa[j + 1], &A_(r, ldr, j + 1, j));
      wa[j] = (wa[j] - sum2) / sdiag[j];
   }
   for (j = 0; j < n; j++) {
      l = ipvt[j];
      x[l] = wa[j];
   }
}

/* minpack.f95:3617-3920 */
static void lmpar(int n, double *r, int ldr, const int *ipvt, const double *diag, const double *qtb, double delta,
                  double *par, double *x, double *sdiag)
{
   const double dwarf = SO_DWARF;   /* tiny(dwarf) */
   double wa1[SO_MP_MAXN], wa2[SO_MP_MAXN], dxnorm, fp, gnorm, parc, parl, paru, sum2, temp, sq;
   int i, iter, j, k, l, nsing;
   nsing = n;
   for (j = 0; j < n; j++) {
      wa1[j] = qtb[j];
      if (A_(r, ldr, j, j) == 0.0 && nsing == n) nsing = j;
      if (nsing < n) wa1[j] = 0.0;
   }
   for (k = 1; k <= nsing; k++) {
      j = nsing - k;
      wa1[j] = wa1[j] / A_(r, ldr, j, j);
      temp = wa1[j];
      for (i = 0; i < j; i++) wa1[i] = wa1[i] - A_(r, ldr, i, j) * temp;
   }
   for (j = 0; j < n; j++) {
      l = ipvt[j];
      x[l] = wa1[j];
   }
   iter = 0;
   for (j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
   dxnorm = enorm(n, wa2);
   fp = dxnorm - delta;
   if (fp <= 0.1 * delta) {
      *par = 0.0;
      return;
   }
   parl = 0.0;
   if (n <= nsing) {
      for (j = 0; j < n; j++) {
         l = ipvt[j];
         wa1[j] = diag[l] * (wa2[l] / dxnorm);
      }
      for (j = 0; j < n; j++) {
         sum2 = dotp(j, wa1, &A_(r, ldr, 0, j));
         wa1[j] = (wa1[j] - sum2) / A_(r, ldr, j, j);
      }
      temp = enorm(n, wa1);
      parl = ((fp / delta) / temp) / temp;
   }
   for (j = 0; j < n; j++) {
      sum2 = dotp(j + 1, qtb, &A_(r, ldr, 0, j));
      l = ipvt[j];
      wa1[j] = sum2 / diag[l];
   }
   gnorm = enorm(n, wa1);
   paru = gnorm / delta;
   if (paru == 0.0) paru = dwarf / (delta < 0.1 ? delta : 0.1);
   *par = (*par > parl) ? *par : parl;
   *par = (*par < paru) ? *par : paru;
   if (*par == 0.0) *par = gnorm / dxnorm;
   for (;;) {
      iter = iter + 1;
      if (*par == 0.0) *par = (dwarf > 0.001 * paru) ? dwarf : 0.001 * paru;
      sq = sqrt(*par);
      for (j = 0; j < n; j++) wa1[j] = sq * diag[j];
      qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag);
      for (j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
      dxnorm = enorm(n, wa2);
      temp = fp;
      fp = dxnorm - delta;
      if (fabs(fp) <= 0.1 * delta) break;
      if (parl == 0.0 && fp <= temp && temp < 0.0) break;
      else if (iter == 10) break;
      for (j = 0; j < n; j++) {
         l = ipvt[j];
         wa1[j] = diag[l] * (wa2[l] / dxnorm);
      }
      for (j = 0; j < n; j++) {
         wa1[j] = wa1[j] / sdiag[j];
         temp = wa1[j];
         for (i = j + 1; i < n; i++) wa1[i] = wa1[i] - A_(r, ldr, i, j) * temp;
      }
      temp = enorm(n, wa1);
      parc = ((fp / delta) / temp) / temp;
      if (0.0 < fp) parl = (parl > *par) ? parl : *par;
      else if (fp < 0.0) paru = (paru < *par) ? paru : *par;
      *par = (parl > *par + parc) ? parl : *par + parc;
   }
   /* (iter > 0 here: the `if (iter == 0) par = 0` of the reference cannot fire) */
}

/* minpack.f95:3477-3616 (lmdif1) + :2942-3476 (lmdif) + :740-860 (fdjac2).  x[n] in/out, fvec[m] out; returns info. */
int so_lmdif1(so_lm_fcn fcn, void *ctx, int m, int n, double *x, double *fvec, double tol)
{
   const double epsmch = SO_EPSMCH;
   const double factor = 100.0, ftol = tol, xtol = tol, gtol = 0.0, epsfcn = 0.0;
   const int maxfev = 200 * (n + 1), ldfjac = m;
   double fjac[SO_MP_MAXM * SO_MP_MAXN], diag[SO_MP_MAXN], qtf[SO_MP_MAXN], wa1[SO_MP_MAXN], wa2[SO_MP_MAXN],
      wa3[SO_MP_MAXN], wa4[SO_MP_MAXM], wa[SO_MP_MAXM];
   int ipvt[SO_MP_MAXN];
   double actred, delta = 0, dirder, fnorm, fnorm1, gnorm, par, pnorm, prered, ratio, sum2, temp = 0, temp1, temp2,
                                 xnorm = 0, eps, h;
   int i, j, l, iter, info = 0, nfev;
   if (n <= 0 || m < n || tol < 0.0) return 0;
   if (n > SO_MP_MAXN || m > SO_MP_MAXM) return -99;
   fcn(m, n, x, fvec, ctx);
   nfev = 1;
   fnorm = enorm(m, fvec);
   par = 0.0;
   iter = 1;
   for (;;) {   /* 30 */
      /* fdjac2 */
      eps = sqrt(epsfcn > epsmch ? epsfcn : epsmch);
      for (j = 0; j < n; j++) {
         temp = x[j];
         h = eps * fabs(temp);
         if (h == 0.0) h = eps;
         x[j] = temp + h;
         fcn(m, n, x, wa, ctx);
         x[j] = temp;
         for (i = 0; i < m; i++) A_(fjac, ldfjac, i, j) = (wa[i] - fvec[i]) / h;
      }
      nfev = nfev + n;
      qrfac(m, n, fjac, ldfjac, 1, ipvt, wa1, wa2);
      if (iter == 1) {
         for (j = 0; j < n; j++) {
            diag[j] = wa2[j];
            if (wa2[j] == 0.0) diag[j] = 1.0;
         }
         for (j = 0; j < n; j++) wa3[j] = diag[j] * x[j];
         xnorm = enorm(n, wa3);
         delta = factor * xnorm;
         if (delta == 0.0) delta = factor;
      }
      for (i = 0; i < m; i++) wa4[i] = fvec[i];
      for (j = 0; j < n; j++) {
         if (A_(fjac, ldfjac, j, j) != 0.0) {
            sum2 = dotp(m - j, &wa4[j], &A_(fjac, ldfjac, j, j));
            temp = -sum2 / A_(fjac, ldfjac, j, j);
            for (i = j; i < m; i++) wa4[i] = wa4[i] + A_(fjac, ldfjac, i, j) * temp;
         }
         A_(fjac, ldfjac, j, j) = wa1[j];
         qtf[j] = wa4[j];
      }
      gnorm = 0.0;
      if (fnorm != 0.0) {
         for (j = 0; j < n; j++) {
            l = ipvt[j];
            if (wa2[l] != 0.0) {
               double g;
               sum2 = 0.0;
               for (i = 0; i <= j; i++) sum2 = sum2 + A_(fjac, ldfjac, i, j) * (qtf[i] / fnorm);
               g = fabs(sum2 / wa2[l]);
               gnorm = (gnorm > g) ? gnorm : g;
            }
         }
      }
      if (gnorm <= gtol) return 4;
      for (j = 0; j < n; j++) diag[j] = (diag[j] > wa2[j]) ? diag[j] : wa2[j];
      for (;;) {   /* 200 */
         lmpar(n, fjac, ldfjac, ipvt, diag, qtf, delta, &par, wa1, wa2);
         for (j = 0; j < n; j++) wa1[j] = -wa1[j];
         for (j = 0; j < n; j++) wa2[j] = x[j] + wa1[j];
         for (j = 0; j < n; j++) wa3[j] = diag[j] * wa1[j];
         pnorm = enorm(n, wa3);
         if (iter == 1) delta = (delta < pnorm) ? delta : pnorm;
         fcn(m, n, wa2, wa4, ctx);
         nfev = nfev + 1;
         fnorm1 = enorm(m, wa4);
         if (0.1 * fnorm1 < fnorm) {
            double q = fnorm1 / fnorm;
            actred = 1.0 - q * q;
         } else {
            actred = -1.0;
         }
         for (j = 0; j < n; j++) {
            wa3[j] = 0.0;
            l = ipvt[j];
            temp = wa1[l];
            for (i = 0; i <= j; i++) wa3[i] = wa3[i] + A_(fjac, ldfjac, i, j) * temp;
         }
         temp1 = enorm(n, wa3) / fnorm;
         temp2 = (sqrt(par) * pnorm) / fnorm;
         prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
         dirder = -(temp1 * temp1 + temp2 * temp2);
         ratio = 0.0;
         if (prered != 0.0) ratio = actred / prered;
         if (ratio <= 0.25) {
            if (actred >= 0.0) temp = 0.5;
            if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
            if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
            delta = temp * ((delta < pnorm / 0.1) ? delta : pnorm / 0.1);
            par = par / temp;
         } else {
            if (par == 0.0 || ratio >= 0.75) {
               delta = 2.0 * pnorm;
               par = 0.5 * par;
            }
         }
         if (0.0001 <= ratio) {
            for (j = 0; j < n; j++) x[j] = wa2[j];
            for (j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
            for (i = 0; i < m; i++) fvec[i] = wa4[i];
            xnorm = enorm(n, wa2);
            fnorm = fnorm1;
            iter = iter + 1;
         }
         if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
         if (delta <= xtol * xnorm) info = 2;
         if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
         if (info != 0) return info;
         if (maxfev <= nfev) info = 5;
         if (fabs(actred) <= epsmch && prered <= epsmch && 0.5 * ratio <= 1.0) info = 6;
         if (delta <= epsmch * xnorm) info = 7;
         if (gnorm <= epsmch) info = 8;
         if (info != 0) return (info == 8) ? 4 : info;
         if (!(ratio < 0.0001)) break;
      }
   }
}

/* minpack.f95:4642-4742 (square case m == n is all hybrd needs, but the statements are the general ones) */
static void qform(int m, int n, double *q, int ldq)
{
   double wa[SO_MP_MAXN], temp;
   int i, j, k, l, minmn;
   minmn = m < n ? m : n;
   for (j = 1; j < minmn; j++)
      for (i = 0; i < j; i++) A_(q, ldq, i, j) = 0.0;
   for (j = n; j < m; j++) {
      for (i = 0; i < m; i++) A_(q, ldq, i, j) = 0.0;
      A_(q, ldq, j, j) = 1.0;
   }
   for (l = 1; l <= minmn; l++) {
      k = minmn - l;
      for (i = k; i < m; i++) { wa[i] = A_(q, ldq, i, k); A_(q, ldq, i, k) = 0.0; }
      A_(q, ldq, k, k) = 1.0;
      if (wa[k] != 0.0) {
         for (j = k; j < m; j++) {
            temp = dotp(m - k, &wa[k], &A_(q, ldq, k, j)) / wa[k];
            for (i = k; i < m; i++) A_(q, ldq, i, j) = A_(q, ldq, i, j) - temp * wa[i];
         }
      }
   }
}

/* minpack.f95:171-380; r is the upper triangle stored by rows, 0-based here */
static void dogleg(int n, const double *r, const double *diag, const double *qtb, double delta, double *x)
{
   const double epsmch = SO_EPSMCH;
   double wa1[SO_MP_MAXN], wa2[SO_MP_MAXN], alpha, bnorm, gnorm, qnorm, sgnorm, sum2, temp;
   int i, j, jj, k, l;
   jj = (n * (n + 1)) / 2 + 1;   /* 1-based positions, as in the reference */
   for (k = 1; k <= n; k++) {
      j = n - k + 1;
      jj = jj - k;
      l = jj + 1;
      sum2 = 0.0;
      for (i = j + 1; i <= n; i++) {
         sum2 = sum2 + r[l - 1] * x[i - 1];
         l = l + 1;
      }
      temp = r[jj - 1];
      if (temp == 0.0) {
         l = j;
         for (i = 1; i <= j; i++) {
            temp = (temp > fabs(r[l - 1])) ? temp : fabs(r[l - 1]);
            l = l + n - i;
         }
         if (temp == 0.0) temp = epsmch;
         else temp = epsmch * temp;
      }
      x[j - 1] = (qtb[j - 1] - sum2) / temp;
   }
   for (j = 0; j < n; j++) { wa1[j] = 0.0; wa2[j] = diag[j] * x[j]; }
   qnorm = enorm(n, wa2);
   if (qnorm <= delta) return;
   l = 1;
   for (j = 1; j <= n; j++) {
      temp = qtb[j - 1];
      for (i = j; i <= n; i++) {
         wa1[i - 1] = wa1[i - 1] + r[l - 1] * temp;
         l = l + 1;
      }
      wa1[j - 1] = wa1[j - 1] / diag[j - 1];
   }
   gnorm = enorm(n, wa1);
   sgnorm = 0.0;
   alpha = delta / qnorm;
   if (gnorm != 0.0) {
      for (j = 0; j < n; j++) wa1[j] = (wa1[j] / gnorm) / diag[j];
      l = 1;
      for (j = 1; j <= n; j++) {
         sum2 = 0.0;
         for (i = j; i <= n; i++) {
            sum2 = sum2 + r[l - 1] * wa1[i - 1];
            l = l + 1;
         }
         wa2[j - 1] = sum2;
      }
      temp = enorm(n, wa2);
      sgnorm = (gnorm / temp) / temp;
      alpha = 0.0;
      if (sgnorm < delta) {
         double dq = delta / qnorm, sd = sgnorm / delta, t1;
         bnorm = enorm(n, qtb);
         temp = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
         t1 = temp - dq;
         temp = temp - dq * (sd * sd) + sqrt(t1 * t1 + (1.0 - dq * dq) * (1.0 - sd * sd));
         alpha = (dq * (1.0 - sd * sd)) / temp;
      }
   }
   temp = (1.0 - alpha) * ((sgnorm < delta) ? sgnorm : delta);
   for (j = 0; j < n; j++) x[j] = temp * wa1[j] + alpha * x[j];
}

/* minpack.f95:5266-5499, 1-based positions kept */
static void r1updt(int m, int n, double *s, const double *u, double *v, double *w, int *sing)
{
   const double giant = SO_GIANT;
   double cos_, cotan, sin_, tan_, tau, temp;
   int i, j, jj, l;
   jj = (n * (2 * m - n + 1)) / 2 - (m - n);
   l = jj;
   for (i = n; i <= m; i++) {
      w[i - 1] = s[l - 1];
      l = l + 1;
   }
   for (j = n - 1; j >= 1; j--) {
      jj = jj - (m - j + 1);
      w[j - 1] = 0.0;
      if (v[j - 1] != 0.0) {
         if (fabs(v[n - 1]) < fabs(v[j - 1])) {
            cotan = v[n - 1] / v[j - 1];
            sin_ = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
            cos_ = sin_ * cotan;
            tau = 1.0;
            if (fabs(cos_) * giant > 1.0) tau = 1.0 / cos_;
         } else {
            tan_ = v[j - 1] / v[n - 1];
            cos_ = 0.5 / sqrt(0.25 + 0.25 * (tan_ * tan_));
            sin_ = cos_ * tan_;
            tau = sin_;
         }
         v[n - 1] = sin_ * v[j - 1] + cos_ * v[n - 1];
         v[j - 1] = tau;
         l = jj;
         for (i = j; i <= m; i++) {
            temp = cos_ * s[l - 1] - sin_ * w[i - 1];
            w[i - 1] = sin_ * s[l - 1] + cos_ * w[i - 1];
            s[l - 1] = temp;
            l = l + 1;
         }
      }
   }
   for (i = 0; i < m; i++) w[i] = w[i] + v[n - 1] * u[i];
   *sing = 0;
   for (j = 1; j <= n - 1; j++) {
      if (w[j - 1] != 0.0) {
         if (fabs(s[jj - 1]) < fabs(w[j - 1])) {
            cotan = s[jj - 1] / w[j - 1];
            sin_ = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
            cos_ = sin_ * cotan;
            if (1.0 < fabs(cos_) * giant) tau = 1.0 / cos_;
            else tau = 1.0;
         } else {
            tan_ = w[j - 1] / s[jj - 1];
            cos_ = 0.5 / sqrt(0.25 + 0.25 * (tan_ * tan_));
            sin_ = cos_ * tan_;
            tau = sin_;
         }
         l = jj;
         for (i = j; i <= m; i++) {
            temp = cos_ * s[l - 1] + sin_ * w[i - 1];
            w[i - 1] = -sin_ * s[l - 1] + cos_ * w[i - 1];
            s[l - 1] = temp;
            l = l + 1;
         }
         w[j - 1] = tau;
      }
      if (s[jj - 1] == 0.0) *sing = 1;
      jj = jj + (m - j + 1);
   }
   l = jj;
   for (i = n; i <= m; i++) {
      s[l - 1] = w[i - 1];
      l = l + 1;
   }
   if (s[jj - 1] == 0.0) *sing = 1;
}

/* minpack.f95:5155-5265 */
static void r1mpyq(int m, int n, double *a, int lda, const double *v, const double *w)
{
   double c, s, temp;
   int i, j;
   for (j = n - 1; j >= 1; j--) {
      if (1.0 < fabs(v[j - 1])) {
         c = 1.0 / v[j - 1];
         s = sqrt(1.0 - c * c);
      } else {
         s = v[j - 1];
         c = sqrt(1.0 - s * s);
      }
      for (i = 0; i < m; i++) {
         temp = c * A_(a, lda, i, j - 1) - s * A_(a, lda, i, n - 1);
         A_(a, lda, i, n - 1) = s * A_(a, lda, i, j - 1) + c * A_(a, lda, i, n - 1);
         A_(a, lda, i, j - 1) = temp;
      }
   }
   for (j = 1; j <= n - 1; j++) {
      if (fabs(w[j - 1]) > 1.0) {
         c = 1.0 / w[j - 1];
         s = sqrt(1.0 - c * c);
      } else {
         s = w[j - 1];
         c = sqrt(1.0 - s * s);
      }
      for (i = 0; i < m; i++) {
         temp = c * A_(a, lda, i, j - 1) + s * A_(a, lda, i, n - 1);
         A_(a, lda, i, n - 1) = -s * A_(a, lda, i, j - 1) + c * A_(a, lda, i, n - 1);
         A_(a, lda, i, j - 1) = temp;
      }
   }
}

/* minpack.f95:1413-1557 (hybrd1) + :862-1411 (hybrd) + :560-739 (fdjac1).  x[n] in/out, fvec[n] out; returns info. */
int so_hybrd1(so_hy_fcn fcn, void *ctx, int n, double *x, double *fvec, double tol)
{
   const double epsmch = SO_EPSMCH;
   const double xtol = tol, epsfcn = 0.0, factor = 100.0;
   const int maxfev = 200 * (n + 1), ldfjac = n;
   double fjac[SO_MP_MAXN * SO_MP_MAXN], r[(SO_MP_MAXN * (SO_MP_MAXN + 1)) / 2], diag[SO_MP_MAXN], qtf[SO_MP_MAXN],
      wa1[SO_MP_MAXN], wa2[SO_MP_MAXN], wa3[SO_MP_MAXN], wa4[SO_MP_MAXN];
   double actred, delta = 0, fnorm, fnorm1, pnorm, prered, ratio, sum2, temp, xnorm = 0, eps, h;
   int i, j, l, iter, info = 0, nfev, msum, ncsuc, ncfail, nslow1, nslow2, jeval, sing;
   if (n <= 0 || tol < 0.0) return 0;
   if (n > SO_MP_MAXN) return -99;
   for (j = 0; j < n; j++) { diag[j] = 1.0; qtf[j] = 0.0; }   /* mode = 2 */
   for (j = 0; j < n * n; j++) fjac[j] = 0.0;
   for (j = 0; j < (n * (n + 1)) / 2; j++) r[j] = 0.0;
   fcn(n, x, fvec, ctx);
   nfev = 1;
   fnorm = enorm(n, fvec);
   msum = (2 * n - 1 < n) ? 2 * n - 1 : n;   /* min(ml + mu + 1, n) with ml = mu = n - 1 */
   iter = 1;
   ncsuc = 0; ncfail = 0; nslow1 = 0; nslow2 = 0;
   for (;;) {   /* 30 */
      jeval = 1;
      /* fdjac1, dense branch (n <= ml + mu + 1) */
      eps = sqrt(epsfcn > epsmch ? epsfcn : epsmch);
      for (j = 0; j < n; j++) {
         temp = x[j];
         h = eps * fabs(temp);
         if (h == 0.0) h = eps;
         x[j] = temp + h;
         fcn(n, x, wa1, ctx);
         x[j] = temp;
         for (i = 0; i < n; i++) A_(fjac, ldfjac, i, j) = (wa1[i] - fvec[i]) / h;
      }
      nfev = nfev + msum;
      qrfac(n, n, fjac, ldfjac, 0, NULL, wa1, wa2);
      if (iter == 1) {
         for (j = 0; j < n; j++) wa3[j] = diag[j] * x[j];
         xnorm = enorm(n, wa3);
         delta = factor * xnorm;
         if (delta == 0.0) delta = factor;
      }
      for (i = 0; i < n; i++) qtf[i] = fvec[i];
      for (j = 0; j < n; j++) {
         if (A_(fjac, ldfjac, j, j) != 0.0) {
            temp = -dotp(n - j, &qtf[j], &A_(fjac, ldfjac, j, j)) / A_(fjac, ldfjac, j, j);
            for (i = j; i < n; i++) qtf[i] = qtf[i] + A_(fjac, ldfjac, i, j) * temp;
         }
      }
      sing = 0;
      for (j = 1; j <= n; j++) {
         l = j;
         for (i = 1; i <= j - 1; i++) {
            r[l - 1] = A_(fjac, ldfjac, i - 1, j - 1);
            l = l + n - i;
         }
         r[l - 1] = wa1[j - 1];
         if (wa1[j - 1] == 0.0) sing = 1;
      }
      qform(n, n, fjac, ldfjac);
      for (;;) {   /* 180 */
         dogleg(n, r, diag, qtf, delta, wa1);
         for (j = 0; j < n; j++) wa1[j] = -wa1[j];
         for (j = 0; j < n; j++) wa2[j] = x[j] + wa1[j];
         for (j = 0; j < n; j++) wa3[j] = diag[j] * wa1[j];
         pnorm = enorm(n, wa3);
         if (iter == 1) delta = (delta < pnorm) ? delta : pnorm;
         fcn(n, wa2, wa4, ctx);
         nfev = nfev + 1;
         fnorm1 = enorm(n, wa4);
         actred = -1.0;
         if (fnorm1 < fnorm) {
            double q = fnorm1 / fnorm;
            actred = 1.0 - q * q;
         }
         l = 1;
         for (i = 1; i <= n; i++) {
            sum2 = 0.0;
            for (j = i; j <= n; j++) {
               sum2 = sum2 + r[l - 1] * wa1[j - 1];
               l = l + 1;
            }
            wa3[i - 1] = qtf[i - 1] + sum2;
         }
         temp = enorm(n, wa3);
         prered = 0.0;
         if (temp < fnorm) {
            double q = temp / fnorm;
            prered = 1.0 - q * q;
         }
         ratio = 0.0;
         if (0.0 < prered) ratio = actred / prered;
         if (ratio < 0.1) {
            ncsuc = 0;
            ncfail = ncfail + 1;
            delta = 0.5 * delta;
         } else {
            ncfail = 0;
            ncsuc = ncsuc + 1;
            if (0.5 <= ratio || 1 < ncsuc) delta = (delta > pnorm / 0.5) ? delta : pnorm / 0.5;
            if (fabs(ratio - 1.0) <= 0.1) delta = pnorm / 0.5;
         }
         if (0.0001 <= ratio) {
            for (j = 0; j < n; j++) x[j] = wa2[j];
            for (j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
            for (j = 0; j < n; j++) fvec[j] = wa4[j];
            xnorm = enorm(n, wa2);
            fnorm = fnorm1;
            iter = iter + 1;
         }
         nslow1 = nslow1 + 1;
         if (0.001 <= actred) nslow1 = 0;
         if (jeval) nslow2 = nslow2 + 1;
         if (0.1 <= actred) nslow2 = 0;
         if (delta <= xtol * xnorm || fnorm == 0.0) info = 1;
         if (info != 0) return info;
         if (maxfev <= nfev) info = 2;
         {
            double d01 = 0.1 * delta;
            if (0.1 * ((d01 > pnorm) ? d01 : pnorm) <= epsmch * xnorm) info = 3;
         }
         if (nslow2 == 5) info = 4;
         if (nslow1 == 10) info = 5;
         if (info != 0) return (info == 5) ? 4 : info;
         if (ncfail == 2) break;   /* 290 -> 30 */
         for (j = 0; j < n; j++) {
            sum2 = dotp(n, wa4, &A_(fjac, ldfjac, 0, j));
            wa2[j] = (sum2 - wa3[j]) / pnorm;
            wa1[j] = diag[j] * ((diag[j] * wa1[j]) / pnorm);
            if (0.0001 <= ratio) qtf[j] = sum2;
         }
         r1updt(n, n, r, wa1, wa2, wa3, &sing);
         r1mpyq(n, n, fjac, ldfjac, wa2, wa3);
         r1mpyq(1, n, qtf, 1, wa2, wa3);
         jeval = 0;
      }
   }
}
