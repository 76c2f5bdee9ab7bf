// JSON validation layer with the API of the reference's rayverb/config.h: register required and
// optional fields on a ConfigValidator, then run() it on a parsed object.  Same keys, same enum
// strings, same error messages ("key X not found in config object", "invalid value").
#pragma once
#include "rayverb.h"
#include "helpers.h"

#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

/// A single check-and-read step over a JSON value.
struct JsonValidatorBase
{
    virtual ~JsonValidatorBase() {}
    virtual void run (const rvbjson::Value & value) const = 0;
};

/// JsonGetter <T>: check() says whether the value has the right shape, get() reads it.
template <typename T> struct JsonGetter;

template <typename T>
struct ValueJsonValidator: public JsonValidatorBase, public JsonGetter <T>
{
    ValueJsonValidator (T & t): JsonGetter <T> (t) {}
    virtual void run (const rvbjson::Value & value) const
    {
        if (! JsonGetter <T>::check (value))
            throw std::runtime_error ("invalid value");
        JsonGetter <T>::get (value);
    }
};

/// Throws if the key is missing (config.h:104-115).
struct RequiredValidator
{
    RequiredValidator (const std::string & s): key (s) {}
    bool validate (const rvbjson::Value & value) const
    {
        if (! value.HasMember (key))
            throw std::runtime_error ("key " + key + " not found in config object");
        return true;
    }
    const std::string & getString() const { return key; }
    std::string key;
};

/// Skips the field if the key is missing (config.h:120-128).
struct OptionalValidator
{
    OptionalValidator (const std::string & s): key (s) {}
    bool validate (const rvbjson::Value & value) const { return value.HasMember (key); }
    const std::string & getString() const { return key; }
    std::string key;
};

template <typename T, typename U>
struct FieldJsonValidator: public ValueJsonValidator <T>, public U
{
    FieldJsonValidator (const std::string & s, T & t): ValueJsonValidator <T> (t), U (s) {}
    virtual void run (const rvbjson::Value & value) const
    {
        if (U::validate (value))
            ValueJsonValidator <T>::run (value [U::getString()]);
    }
};

/// Register fields, then run() on the config object (config.h:58-81).
class ConfigValidator: public JsonValidatorBase
{
public:
    template <typename T>
    void addOptionalValidator (const std::string & s, T & t)
    {
        validators.emplace_back (new FieldJsonValidator <T, OptionalValidator> (s, t));
    }
    template <typename T>
    void addRequiredValidator (const std::string & s, T & t)
    {
        validators.emplace_back (new FieldJsonValidator <T, RequiredValidator> (s, t));
    }
    virtual void run (const rvbjson::Value & value) const
    {
        for (const auto & i : validators) i->run (value);
    }
private:
    std::vector <std::unique_ptr <JsonValidatorBase>> validators;
};

// ---- getters -----------------------------------------------------------------------------------
template<> struct JsonGetter <double>
{
    JsonGetter (double & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsNumber(); }
    void get (const rvbjson::Value & v) const { t = v.GetDouble(); }
    double & t;
};
template<> struct JsonGetter <float>
{
    JsonGetter (float & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsNumber(); }
    void get (const rvbjson::Value & v) const { t = (float) v.GetDouble(); }
    float & t;
};
template<> struct JsonGetter <bool>
{
    JsonGetter (bool & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsBool(); }
    void get (const rvbjson::Value & v) const { t = v.GetBool(); }
    bool & t;
};
template<> struct JsonGetter <int>
{
    JsonGetter (int & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsInt(); }
    void get (const rvbjson::Value & v) const { t = v.GetInt(); }
    int & t;
};

/// Numeric arrays of a fixed length into cl_floatN.
template <typename T, int LENGTH>
struct JsonArrayGetter
{
    JsonArrayGetter (T & t): t (t) {}
    bool check (const rvbjson::Value & v) const
    {
        if (! v.IsArray() || v.Size() != (size_t) LENGTH) return false;
        for (int i = 0; i != LENGTH; ++i)
            if (! v [i].IsNumber()) return false;
        return true;
    }
    void get (const rvbjson::Value & v) const
    {
        for (int i = 0; i != LENGTH; ++i) t.s [i] = static_cast <cl_float> (v [i].GetDouble());
    }
    T & t;
};
template<> struct JsonGetter <cl_float3>: public JsonArrayGetter <cl_float3, 3>
{
    JsonGetter (cl_float3 & t): JsonArrayGetter (t) { t.s [3] = 0; }
};
template<> struct JsonGetter <cl_float8>: public JsonArrayGetter <cl_float8, 8>
{
    JsonGetter (cl_float8 & t): JsonArrayGetter (t) {}
};

template<> struct JsonGetter <Surface>
{
    JsonGetter (Surface & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsObject(); }
    void get (const rvbjson::Value & v) const
    {
        ConfigValidator cv;
        cv.addRequiredValidator ("specular", t.specular);
        cv.addRequiredValidator ("diffuse", t.diffuse);
        cv.run (v);
    }
    Surface & t;
};

/// Strings mapped to enum values.
template <typename T>
struct JsonEnumGetter
{
    JsonEnumGetter (T & t, const std::map <std::string, T> & m): t (t), stringkeys (m) {}
    bool check (const rvbjson::Value & v) const { return v.IsString() && stringkeys.count (v.GetString()) != 0; }
    void get (const rvbjson::Value & v) const { t = stringkeys.at (v.GetString()); }
    T & t;
    const std::map <std::string, T> stringkeys;
};
template<> struct JsonGetter <RayverbFiltering::FilterType>: public JsonEnumGetter <RayverbFiltering::FilterType>
{
    JsonGetter (RayverbFiltering::FilterType & t)
    :   JsonEnumGetter
        (   t
        ,   {   {"sinc",           RayverbFiltering::FILTER_TYPE_WINDOWED_SINC}
            ,   {"onepass",        RayverbFiltering::FILTER_TYPE_BIQUAD_ONEPASS}
            ,   {"twopass",        RayverbFiltering::FILTER_TYPE_BIQUAD_TWOPASS}
            ,   {"linkwitz_riley", RayverbFiltering::FILTER_TYPE_LINKWITZ_RILEY}
            }
        )
    {}
};
template<> struct JsonGetter <OutputMode>: public JsonEnumGetter <OutputMode>
{
    JsonGetter (OutputMode & t)
    :   JsonEnumGetter (t, {{"all", ALL}, {"image_only", IMAGE_ONLY}, {"diffuse_only", DIFFUSE_ONLY}})
    {}
};

template<> struct JsonGetter <Speaker>
{
    JsonGetter (Speaker & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsObject(); }
    void get (const rvbjson::Value & v) const
    {
        ConfigValidator cv;
        cv.addRequiredValidator ("direction", t.direction);
        cv.addRequiredValidator ("shape", t.coefficient);
        cv.run (v);
    }
    Speaker & t;
};

template<> struct JsonGetter <HrtfConfig>
{
    JsonGetter (HrtfConfig & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsObject(); }
    void get (const rvbjson::Value & v) const
    {
        ConfigValidator cv;
        cv.addRequiredValidator ("facing", t.facing);
        cv.addRequiredValidator ("up", t.up);
        cv.run (v);
        unit (t.facing);   // HRTF vectors are normalised at parse time (config.h:395-408);
        unit (t.up);       // speaker directions are not (the kernel normalises them)
    }
    HrtfConfig & t;
private:
    static void unit (cl_float3 & v)
    {
        const cl_float len = (cl_float) (1.0 / std::sqrt (v.s [0] * v.s [0] + v.s [1] * v.s [1] + v.s [2] * v.s [2]));
        for (int i = 0; i != 4; ++i) v.s [i] *= len;
    }
};

template <typename T>
struct JsonGetter <std::vector <T>>
{
    JsonGetter (std::vector <T> & t): t (t) {}
    bool check (const rvbjson::Value & v) const { return v.IsArray(); }
    void get (const rvbjson::Value & v) const
    {
        for (size_t i = 0; i != v.Size(); ++i)
        {
            T temp = T();
            ValueJsonValidator <T> getter (temp);
            getter.run (v [i]);
            t.push_back (temp);
        }
    }
    std::vector <T> & t;
};

/// Exactly one of "speakers" / "hrtf" must be present (config.h:433-470).
template<> struct JsonGetter <AttenuationModel>
{
    JsonGetter (AttenuationModel & t): t (t) {}
    bool check (const rvbjson::Value & v) const
    {
        return v.IsObject() && (v.HasMember ("speakers") ? 1 : 0) + (v.HasMember ("hrtf") ? 1 : 0) == 1;
    }
    void get (const rvbjson::Value & v) const
    {
        ConfigValidator cv;
        if (v.HasMember ("speakers"))
        {
            t.mode = AttenuationModel::SPEAKER;
            cv.addRequiredValidator ("speakers", t.speakers);
        }
        if (v.HasMember ("hrtf"))
        {
            t.mode = AttenuationModel::HRTF;
            cv.addRequiredValidator ("hrtf", t.hrtf);
        }
        cv.run (v);
    }
    AttenuationModel & t;
};
