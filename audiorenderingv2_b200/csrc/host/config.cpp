// config.cpp -- config.json loader of libarv2.
//
// Keeps the schema, defaults and rounding quirks of Context::loadContext
// (OR/Context.cpp:15-165): every key optional; ir_length_in_seconds, width, height,
// re_render_distance_threshold, re_render_angle_threshold, ray_max_bounces and
// hrtf_absorption_rate are round()ed (half away from zero) when present -- so the
// shipped "hrtf_absorption_rate": 0.9 becomes 1.0 (Context.cpp:143-145);
// ray_distance_threshold and materials_file_path are accepted and ignored.
// The reference parses with the vendored cJSON; this is a small own parser.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

#include "../arv2_internal.h"

namespace arv2 {

namespace {

struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue* get(const char* key) const
    {
        if (kind != Obj) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;   // first match, like cJSON_GetObjectItem
        return nullptr;
    }
};

struct Parser {
    const char* s; const char* e; std::string err;
    void ws() { while (s < e && (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r')) ++s; }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    bool value(JValue& v, int depth)
    {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (s >= e) return fail("unexpected end");
        if (*s == '{') {
            v.kind = JValue::Obj; ++s; ws();
            if (s < e && *s == '}') { ++s; return true; }
            for (;;) {
                ws();
                JValue k;
                if (s >= e || *s != '"' || !string(k.str)) return fail("expected key");
                ws();
                if (s >= e || *s != ':') return fail("expected ':'");
                ++s;
                JValue child;
                if (!value(child, depth + 1)) return false;
                v.obj.emplace_back(k.str, std::move(child));
                ws();
                if (s < e && *s == ',') { ++s; continue; }
                if (s < e && *s == '}') { ++s; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*s == '[') {
            v.kind = JValue::Arr; ++s; ws();
            if (s < e && *s == ']') { ++s; return true; }
            for (;;) {
                JValue child;
                if (!value(child, depth + 1)) return false;
                v.arr.push_back(std::move(child));
                ws();
                if (s < e && *s == ',') { ++s; continue; }
                if (s < e && *s == ']') { ++s; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*s == '"') { v.kind = JValue::Str; return string(v.str); }
        if (e - s >= 4 && !std::strncmp(s, "true", 4)) { v.kind = JValue::Bool; v.b = true; s += 4; return true; }
        if (e - s >= 5 && !std::strncmp(s, "false", 5)) { v.kind = JValue::Bool; v.b = false; s += 5; return true; }
        if (e - s >= 4 && !std::strncmp(s, "null", 4)) { v.kind = JValue::Null; s += 4; return true; }
        char* end = nullptr;
        const double d = std::strtod(s, &end);
        if (end == s) return fail("unexpected token");
        v.kind = JValue::Num; v.num = d; s = end;
        return true;
    }
    bool string(std::string& out)
    {
        ++s;
        while (s < e && *s != '"') {
            if (*s == '\\' && s + 1 < e) {
                ++s;
                switch (*s) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': {
                    if (e - s < 5) return fail("bad \\u escape");
                    unsigned cp = 0;
                    for (int i = 1; i <= 4; ++i) {
                        const char c = s[i];
                        cp <<= 4;
                        if (c >= '0' && c <= '9') cp |= c - '0';
                        else if (c >= 'a' && c <= 'f') cp |= c - 'a' + 10;
                        else if (c >= 'A' && c <= 'F') cp |= c - 'A' + 10;
                        else return fail("bad \\u escape");
                    }
                    s += 4;
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                    else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: out += *s;
                }
                ++s;
            } else out += *s++;
        }
        if (s >= e) return fail("unterminated string");
        ++s;
        return true;
    }
};

void copy_str(char* dst, size_t cap, const std::string& s)
{
    std::snprintf(dst, cap, "%s", s.c_str());
}

bool num(const JValue* v, double* d) { if (v && v->kind == JValue::Num) { *d = v->num; return true; } return false; }
bool vec3(const JValue* v, float out[3])
{
    if (!v || v->kind != JValue::Obj) return false;
    double x, y, z;
    if (num(v->get("x"), &x) && num(v->get("y"), &y) && num(v->get("z"), &z)) {
        out[0] = (float)x; out[1] = (float)y; out[2] = (float)z;
        return true;
    }
    return false;
}

} // namespace

int parse_config(const std::string& json, arv2_config* c, std::string* err)
{
    Parser p{json.data(), json.data() + json.size(), {}};
    JValue root;
    if (!p.value(root, 0)) { *err = "config.json: " + p.err; return ARV2_ERR_IO; }

    std::memset(c, 0, sizeof(*c));
    // defaults, Context.cpp:20-27, 66-71, 113-118
    c->initial_volume = 1.0f; c->ir_length_in_seconds = 2; c->width = 1366; c->height = 768;
    c->re_render_distance_threshold = 3.f; c->re_render_angle_threshold = 5.f;
    copy_str(c->scene_file_path, sizeof c->scene_file_path, "../../assets/models/1D_U.obj");
    c->initial_receiver_pos[0] = -2.5f; c->initial_receiver_pos[1] = 10.0f; c->initial_receiver_pos[2] = 0.0f;
    c->base_power = 100.f; c->rays[0] = c->rays[1] = c->rays[2] = 100.f;
    c->ray_energy_threshold = 0.f; c->ray_max_bounces = 10; c->hrtf_absorption_rate = 0.9f;
    c->seed = 1; c->bands = 1;

    double d;
    if (const JValue* r = root.get("renderer_parameters"); r && r->kind == JValue::Obj) {
        if (num(r->get("initial_volume"), &d)) c->initial_volume = (float)d;
        if (num(r->get("ir_length_in_seconds"), &d)) c->ir_length_in_seconds = (uint32_t)std::round(d);
        if (num(r->get("width"), &d)) c->width = (uint32_t)std::round(d);
        if (num(r->get("height"), &d)) c->height = (uint32_t)std::round(d);
        if (const JValue* b = r->get("write_first_ir_to_file"); b && b->kind == JValue::Bool) c->write_first_ir_to_file = b->b;
        if (const JValue* b = r->get("write_first_output_to_file"); b && b->kind == JValue::Bool) c->write_first_output_to_file = b->b;
        if (num(r->get("re_render_distance_threshold"), &d)) c->re_render_distance_threshold = (float)std::round(d);
        if (num(r->get("re_render_angle_threshold"), &d)) c->re_render_angle_threshold = (float)std::round(d);
    }
    if (const JValue* s = root.get("scene_parameters"); s && s->kind == JValue::Obj) {
        if (const JValue* b = s->get("mono"); b && b->kind == JValue::Bool) c->mono = b->b;
        if (const JValue* v = s->get("scene_file_path"); v && v->kind == JValue::Str) copy_str(c->scene_file_path, sizeof c->scene_file_path, v->str);
        if (const JValue* v = s->get("audio_file_path"); v && v->kind == JValue::Str) copy_str(c->audio_file_path, sizeof c->audio_file_path, v->str);
        if (const JValue* v = s->get("materials_file_path"); v && v->kind == JValue::Str) copy_str(c->materials_file_path, sizeof c->materials_file_path, v->str);
        vec3(s->get("initial_receiver_pos"), c->initial_receiver_pos);
        vec3(s->get("initial_emitter_pos"), c->initial_emitter_pos);
    }
    if (const JValue* t = root.get("pathtracer_parameters"); t && t->kind == JValue::Obj) {
        if (num(t->get("base_power"), &d)) c->base_power = (float)d;
        vec3(t->get("rays"), c->rays);
        if (num(t->get("ray_energy_threshold"), &d)) c->ray_energy_threshold = (float)d;
        if (num(t->get("ray_max_bounces"), &d)) c->ray_max_bounces = (uint32_t)std::round(d);
        if (num(t->get("hrtf_absorption_rate"), &d)) c->hrtf_absorption_rate = (float)std::round(d);
        if (num(t->get("seed"), &d)) c->seed = (uint64_t)d;
        if (num(t->get("bands"), &d)) c->bands = (int32_t)std::round(d);
        if (const JValue* ms = t->get("materials"); ms && ms->kind == JValue::Arr) {
            for (const JValue& m : ms->arr) {
                const JValue* nm = m.get("name");
                double a;
                if (nm && nm->kind == JValue::Str && num(m.get("mat_absorption"), &a) &&
                    c->n_materials < ARV2_MAX_CONFIG_MATERIALS) {
                    copy_str(c->material_names[c->n_materials], sizeof c->material_names[0], nm->str);
                    c->material_absorption[c->n_materials] = (float)a;
                    c->n_materials++;
                }
            }
        }
    }
    return ARV2_OK;
}

} // namespace arv2
