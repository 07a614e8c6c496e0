/* minijson.c -- see minijson.h */
#include "minijson.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void minijson_init(MiniJson *j, const char *buf, size_t len) {
    j->p = buf;
    j->end = buf + len;
    j->error[0] = '\0';
}

static void ws(MiniJson *j) {
    while (j->p < j->end && isspace((unsigned char) *j->p)) j->p++;
}

static int fail(MiniJson *j, const char *what) {
    snprintf(j->error, sizeof(j->error), "%s near offset %ld", what, (long) (j->end - j->p));
    return -1;
}

static int expect(MiniJson *j, char ch) {
    ws(j);
    if (j->p >= j->end || *j->p != ch) {
        snprintf(j->error, sizeof(j->error), "expected '%c' with %ld bytes left", ch, (long) (j->end - j->p));
        return -1;
    }
    j->p++;
    return 0;
}

static int string(MiniJson *j, char *out, size_t cap) {
    if (expect(j, '"') != 0) return -1;
    size_t k = 0;
    while (j->p < j->end && *j->p != '"') {
        char ch = *j->p++;
        if (ch == '\\' && j->p < j->end) ch = *j->p++;
        if (k + 1 < cap) out[k++] = ch;
    }
    if (j->p >= j->end) return fail(j, "unterminated string");
    j->p++;
    out[k] = '\0';
    return 0;
}

int minijson_number(MiniJson *j, double *out) {
    ws(j);
    char tmp[64];
    size_t k = 0;
    while (j->p < j->end && k + 1 < sizeof(tmp) && (isdigit((unsigned char) *j->p) || strchr("+-.eE", *j->p) != NULL)) tmp[k++] = *j->p++;
    tmp[k] = '\0';
    char *endp = NULL;
    const double v = strtod(tmp, &endp);
    if (k == 0 || endp == tmp || *endp != '\0') return fail(j, "expected a number");
    *out = v;
    return 0;
}

int minijson_bool(MiniJson *j, int *out) {
    ws(j);
    if (j->end - j->p >= 4 && strncmp(j->p, "true", 4) == 0) {
        j->p += 4;
        *out = 1;
        return 0;
    }
    if (j->end - j->p >= 5 && strncmp(j->p, "false", 5) == 0) {
        j->p += 5;
        *out = 0;
        return 0;
    }
    double v;
    if (minijson_number(j, &v) != 0) return fail(j, "expected true, false, 0 or 1");
    *out = v != 0.0;
    return 0;
}

int minijson_number_array(MiniJson *j, double *out, int64_t n) {
    if (expect(j, '[') != 0) return -1;
    for (int64_t i = 0; i < n; i++) {
        if (i > 0 && expect(j, ',') != 0) return -1;
        if (minijson_number(j, &out[i]) != 0) return -1;
    }
    return expect(j, ']');
}

int minijson_skip_value(MiniJson *j) {
    ws(j);
    if (j->p >= j->end) return fail(j, "missing value");
    if (*j->p == '"') {
        char tmp[8];
        return string(j, tmp, sizeof(tmp));
    }
    if (*j->p == '[' || *j->p == '{') {
        const char open = *j->p, close = open == '[' ? ']' : '}';
        int depth = 0;
        while (j->p < j->end) {
            const char ch = *j->p;
            if (ch == '"') {
                char tmp[8];
                if (string(j, tmp, sizeof(tmp)) != 0) return -1;
                continue;
            }
            j->p++;
            if (ch == open) depth++;
            else if (ch == close && --depth == 0) return 0;
        }
        return fail(j, "unterminated array or object");
    }
    while (j->p < j->end && *j->p != ',' && *j->p != '}' && *j->p != ']' && !isspace((unsigned char) *j->p)) j->p++;
    return 0;
}

int minijson_object(MiniJson *j, MiniJsonMember member, void *extra) {
    if (expect(j, '{') != 0) return -1;
    ws(j);
    if (j->p < j->end && *j->p == '}') {
        j->p++;
        return 0;
    }
    for (;;) {
        char key[64];
        ws(j);
        if (string(j, key, sizeof(key)) != 0) return -1;
        if (expect(j, ':') != 0) return -1;
        if (member(j, key, extra) != 0) {
            if (j->error[0] == '\0') snprintf(j->error, sizeof(j->error), "bad value for key %s", key);
            return -1;
        }
        ws(j);
        if (j->p < j->end && *j->p == ',') {
            j->p++;
            continue;
        }
        return expect(j, '}');
    }
}
